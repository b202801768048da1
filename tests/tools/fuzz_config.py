"""Differential fuzzing of freesasa_classifier_from_file() against the compiled reference (inputs on which the reference's
own assert()s abort are filtered out).  usage: python tests/tools/fuzz_config.py SEED N

CPU only.  Round 1: 13 500 + 17 000 + 36 000 cases, no divergence."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from freesasa_b200 import structure as st
from oracle import bindings as ob
seed=int(sys.argv[1]); n=int(sys.argv[2])
mine=st.api(); ref=st.StructureAPI(ob.ref_lib(), ob.RefResult, ob.RefParameters)
mine.lib.freesasa_set_verbosity(2); ref.lib.freesasa_set_verbosity(2)
rng=np.random.default_rng(seed)
base=bytearray(b"name: fuzz\n\ntypes:\nA 1.0 polar # c\nB 2.0 apolar\nC3 1.55 Polar\n# c\n\natoms:\nAA aa A # c\nBB bb B\nANY cc C3\nAA bb B\nCC aa A\n")
alphabet=b" \n\t#:.0123456789ABCabcnametypsol-"
for trial in range(n):
    text=bytearray(base)
    for _ in range(int(rng.integers(1,5))):
        pos=int(rng.integers(0,len(text))); kind=int(rng.integers(0,3))
        if kind==0: text[pos]=alphabet[int(rng.integers(0,len(alphabet)))]
        elif kind==1: del text[pos:pos+int(rng.integers(1,12))]
        else: text[pos:pos]=bytes(alphabet[int(k)] for k in rng.integers(0,len(alphabet),size=int(rng.integers(1,10))))
    t=bytes(text)
    bad=False
    for line in t.split(b"\n"):
        vis=line.split(b"#")[0]
        hits=[k for k in (b"name:",b"types:",b"atoms:") if k in vis]
        if not hits: continue
        sv=vis.strip(b" \t")
        if len(hits)>1 or vis.count(hits[0])>1 or not sv.startswith(hits[0]) or (len(sv)>len(hits[0]) and sv[len(hits[0]):len(hits[0])+1] not in (b" ",b"\t")) or (hits[0]+b"#") in line: bad=True
    if bad or len(max(t.split(b"\n"),key=len))>250: continue
    open('/tmp/fuzzcfg_last_%d.bin'%seed,'wb').write(t)
    cm=mine.classifier_from_text(t); cr=ref.classifier_from_text(t)
    res=[]
    for api,c in ((mine,cm),(ref,cr)):
        if not c: res.append(None); continue
        out=[]
        for r in (b"AA",b"BB",b"CC",b"ANY",b"ZZ"):
            for a in (b"aa",b"bb",b"cc",b"zz"):
                out.append((api.lib.freesasa_classifier_radius(c,r,a), api.lib.freesasa_classifier_class(c,r,a)))
        res.append(out)
    if res[0]!=res[1]:
        h=hashlib.md5(t).hexdigest()[:8]; open('/tmp/fuzzcfg_div_%s.bin'%h,'wb').write(t); print('DIVERGENCE',h, res[0] is None, res[1] is None, flush=True)
print('done',seed)

"""Differential fuzzing of the selection language against the reference's flex/bison parser: random token sequences (valid
and invalid); commands on which the reference dereferences NULL (open resi ranges with identifiers) are filtered out.
usage: python tests/tools/fuzz_select.py SEED N

CPU only.  Round 1: 13 500 + 17 000 + 36 000 cases, no divergence."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, ctypes
from freesasa_b200 import structure as st, workloads as w
from oracle import bindings as ob
from tests.test_select import Selector
seed=int(sys.argv[1]); n=int(sys.argv[2])
mine=st.api(); ref=st.StructureAPI(ob.ref_lib(), ob.RefResult, ob.RefParameters)
text=(w.pdb_text(300, seed=17, chains=3, hetatm=3, unknown=0.05)+w.pdb_atom_line(9001,"CA","ALA","D",-5,1.0,2.0,3.0,"C")+"\n"+w.pdb_atom_line(9003,"SE","MSE","D",8,7.0,2.0,3.0,"SE",icode="A")+"\n").encode()
world=[]
for api in (mine,ref):
    api.lib.freesasa_set_verbosity(2)
    s=api.from_pdb(text,None,st.INCLUDE_HETATM); tree=st.TreeAPI(api)
    rng0=np.random.default_rng(3); result,keep=tree.make_result(rng0.uniform(0,40,size=s.n))
    world.append((Selector(api),s,result,keep))
rng=np.random.default_rng(seed)
toks=["resn","resi","name","symbol","chain","and","or","not","&","|","!","(",")","+","-","\\-",",","ALA","gly","CA","C5'","O","N","1","20","8A","A","B","C","SE","x_1","s","  ","AND","Resn"]
for trial in range(n):
    k=int(rng.integers(1,10))
    parts=[toks[int(i)] for i in rng.integers(0,len(toks),size=k)]
    sep=" " if rng.random()<0.8 else ""
    cmd="s, "+sep.join(parts) if rng.random()<0.85 else sep.join(parts)
    # the reference dereferences NULL for an open-left resi range with a non-number (src/selection.c:466-469)
    import re
    if re.search(r"resi\s*(\S+\s*\+\s*)*-\s*[A-Za-z_]", cmd, re.I) or re.search(r"-\s*(?![0-9\\])\S*[A-Za-z_']", cmd) and "resi" in cmd.lower(): continue
    if "resi" in cmd.lower() and re.search(r"[A-Za-z_']\s*-\s*(\+|\)|and|or|&|\||$|,|!|not|\()", cmd, re.I): continue
    open('/tmp/fuzzsel_last_%d.txt'%seed,'w').write(cmd)
    c=cmd.encode()
    a=world[0][0].run(c,world[0][1],world[0][2]); b=world[1][0].run(c,world[1][1],world[1][2])
    if a!=b: print('DIVERGENCE', repr(cmd), a, b, flush=True)
print('done',seed)

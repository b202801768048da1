"""Differential fuzzing of freesasa_structure_add_atom[_wopt]() sequences (atoms added one by one, classifiers and options
changing from call to call) and freesasa_structure_get_chains() against the compiled reference.
usage: python tests/tools/fuzz_add_atom.py SEED N      CPU only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from freesasa_b200 import structure as st  # noqa: E402
from oracle import bindings as ob  # noqa: E402
from tests.test_ingest import snapshot  # noqa: E402

seed, n = int(sys.argv[1]), int(sys.argv[2])
mine = st.api()
ref = st.StructureAPI(ob.ref_lib(), ob.RefResult, ob.RefParameters)
for api in (mine, ref):
    api.lib.freesasa_set_verbosity(2)
rng = np.random.default_rng(seed)
names = [b" CA ", b" N  ", b" O  ", b" CB ", b" OXT", b"FE  ", b" XX ", b"1HB ", b"HG11", b" SE ", b" C5'", b"ABCD", b" H  "]
residues = [b"ALA", b"GLY", b"MSE", b"HOH", b"ZZZ", b"  A", b" DA", b"UNK"]
numbers = [b"   1 ", b"   1A", b"  -5 ", b"  12 ", b"1234 ", b"   2 ", b"   1"]
chains = [b"A", b"B", b"C", b" ", b"1"]
options = [0, st.SKIP_UNKNOWN, st.HALT_AT_UNKNOWN, st.SKIP_UNKNOWN | st.HALT_AT_UNKNOWN, st.RADIUS_FROM_OCCUPANCY]
for trial in range(n):
    k = int(rng.integers(1, 25))
    plan = [(names[int(rng.integers(0, len(names)))], residues[int(rng.integers(0, len(residues)))],
             numbers[int(rng.integers(0, len(numbers)))], chains[int(rng.integers(0, len(chains)))],
             float(rng.normal()), float(rng.normal()), float(rng.normal()),
             [None, "oons", "naccess", "protor"][int(rng.integers(0, 4))], options[int(rng.integers(0, len(options)))],
             bool(rng.random() < 0.3)) for _ in range(k)]
    group = bytes(rng.choice([65, 66, 67, 32, 49], size=int(rng.integers(1, 4))).tolist())
    out = []
    for api in (mine, ref):
        s = api.new()
        rcs = []
        for name, res, num, ch, x, y, z, cls, opt, plain in plan:
            if plain:
                rcs.append(s.add_atom(name, res, num, ch, x, y, z))
            else:
                rcs.append(s.add_atom(name, res, num, ch, x, y, z, api.classifier(cls) if cls else None, opt))
        snap = snapshot(s) if s.n else None
        sub = s.get_chains(group) if s.n else None
        out.append((rcs, snap, snapshot(sub) if sub else None))
    if out[0] != out[1]:
        print("DIVERGENCE", seed, trial, plan, group, flush=True)
print("done", seed)

"""Differential fuzzing of the result tree against the compiled reference: mutated PDB files (odd chain / residue
arrangements, alternate locations, hetero atoms) -> freesasa_tree_init() in both libraries with the same per-atom values ->
the flattened trees must be equal bit for bit.  usage: python tests/tools/fuzz_tree.py SEED N

CPU only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from freesasa_b200 import structure as st, workloads as w  # noqa: E402
from oracle import bindings as ob  # noqa: E402

seed, n = int(sys.argv[1]), int(sys.argv[2])
mine = st.api()
ref = st.StructureAPI(ob.ref_lib(), ob.RefResult, ob.RefParameters)
sides = []
for api in (mine, ref):
    api.lib.freesasa_set_verbosity(2)
    sides.append((api, st.TreeAPI(api)))
rng = np.random.default_rng(seed)
base = w.pdb_text(80, seed=seed % 5, chains=3, hydrogens=0.1, hetatm=3, altloc=0.1, unknown=0.1).encode().split(b"\n")
chains = b"ABCab 1"
for trial in range(n):
    lines = list(base)
    for _ in range(int(rng.integers(1, 8))):
        k = int(rng.integers(0, len(lines)))
        kind = int(rng.integers(0, 4))
        ln = bytearray(lines[k])
        if kind == 0 and len(ln) > 30:      # another chain label
            ln[21] = chains[int(rng.integers(0, len(chains)))]
        elif kind == 1 and len(ln) > 30:    # another residue number / insertion code
            ln[22:27] = b"%4d%s" % (int(rng.integers(-3, 30)), rng.choice([b" ", b"A", b"B"]))
        elif kind == 2 and len(ln) > 30:    # another residue name
            ln[17:20] = rng.choice([b"ALA", b"GLY", b"HOH", b"  A", b"UNK", b"MSE"])
        else:                               # move the line
            lines.insert(int(rng.integers(0, len(lines))), lines.pop(k))
            continue
        lines[k] = bytes(ln)
    text = b"\n".join(lines)
    options = [0, st.INCLUDE_HETATM, st.INCLUDE_HETATM | st.INCLUDE_HYDROGEN, st.SKIP_UNKNOWN][trial % 4]
    walks = []
    for api, tree in sides:
        s = api.from_pdb(text, None, options)
        if s is None:
            walks.append(None)
            continue
        values = np.random.default_rng(trial).uniform(0, 30, size=s.n)
        result, keep = tree.make_result(values)
        root = tree.init(result, s, b"f")
        walks.append(tree.walk(root))
        tree.free(root)
    if walks[0] != walks[1]:
        open("/tmp/fuzz_tree_div_%d_%d.pdb" % (seed, trial), "wb").write(text)
        print("DIVERGENCE", seed, trial, options, flush=True)
print("done", seed)

"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the dev container only (needs /root/reference and oracle/_ref built by `make -C oracle ref`):

    python tests/golden/make_golden.py

Outputs (committed):
  tests/golden/pdb_fixtures.npz   per structure: coordinates + ProtOr radii exactly as the reference
                                  reads the PDB file (freesasa_structure_from_pdb(f, NULL, 0),
                                  reference src/structure.c:838), and the reference's per-atom SASA
                                  for Lee-Richards n_slices=20 and Shrake-Rupley n_points=100 at 1 thread
  tests/golden/synthetic.npz      seeded synthetic globule + reference per-atom SASA at the benchmark
                                  resolutions (LR 20/100, SR 100/1000)
  tests/golden/totals.json        totals: those published in the reference's own tests
                                  (tests/test_freesasa.c) and those measured here from the reference
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as ob  # noqa: E402
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("workloads", os.path.join(ROOT, "freesasa_b200", "workloads.py"))
workloads = importlib.util.module_from_spec(spec)
spec.loader.exec_module(workloads)

DATA = "/root/reference/tests/data"
HERE = os.path.dirname(os.path.abspath(__file__))
PDBS = ["1ubq", "2jo4", "3bkr", "5dx9", "3bzd_trimmed", "1d3z"]
ALL_PDBS = ["1a0q", "1sui", "2isk", "2jo4", "3bkr", "3bzd_trimmed", "3gnn", "5dx9", "5hdn", "1d3z", "1ubq"]

# Totals asserted by the reference's own unit tests (abs tol 1e-5 there).
PUBLISHED = {
    "1ubq": {"lr20": 4804.055641, "sr100": 4834.716265, "src": "tests/test_freesasa.c:155-178"},
    "3bzd_trimmed": {"sr100": 16133.867124, "src": "tests/test_freesasa.c:302-332"},
    "1d3z": {"sr100": 5000.340175, "src": "tests/test_freesasa.c:432-473"},
}


def main():
    ob.ref_lib().freesasa_set_verbosity(1)  # no warnings about unknown atoms
    out, totals = {}, {"published": PUBLISHED, "measured": {}}
    for name in ALL_PDBS:
        xyz, rad = ob.ref_structure_from_pdb(os.path.join(DATA, name + ".pdb"))
        lr = ob.ref_calc(xyz, rad, ob.LEE_RICHARDS, 1.4, 20, 1)
        sr = ob.ref_calc(xyz, rad, ob.SHRAKE_RUPLEY, 1.4, 100, 1)
        totals["measured"][name] = {"n_atoms": int(len(rad)), "lr20": float(lr.sum()), "sr100": float(sr.sum())}
        if name in PDBS:
            out[name + "_xyz"], out[name + "_radii"] = xyz, rad
            out[name + "_lr20"], out[name + "_sr100"] = lr, sr
        print(name, len(rad), lr.sum(), sr.sum())
    for name, pub in PUBLISHED.items():
        for k in ("lr20", "sr100"):
            if k in pub:
                assert abs(pub[k] - totals["measured"][name][k]) < 1e-5, (name, k)
    np.savez_compressed(os.path.join(HERE, "pdb_fixtures.npz"), **out)

    syn = {}
    xyz, rad = workloads.globule(3000, seed=7)
    syn["g3000_xyz"], syn["g3000_radii"] = xyz, rad
    for key, alg, res in [("lr20", 0, 20), ("lr100", 0, 100), ("sr100", 1, 100), ("sr1000", 1, 1000)]:
        syn["g3000_" + key] = ob.ref_calc(xyz, rad, alg, 1.4, res, 1)
    # far-from-origin copy: PDB files routinely sit hundreds of Å away from the origin
    xyz_off, rad_off = workloads.globule(1500, seed=11, offset=(512.25, -377.5, 941.125))
    syn["off1500_xyz"], syn["off1500_radii"] = xyz_off, rad_off
    syn["off1500_lr20"] = ob.ref_calc(xyz_off, rad_off, 0, 1.4, 20, 1)
    syn["off1500_sr100"] = ob.ref_calc(xyz_off, rad_off, 1, 1.4, 100, 1)
    np.savez_compressed(os.path.join(HERE, "synthetic.npz"), **syn)

    big, brad = workloads.globule(100000)
    totals["measured"]["globule100k"] = {
        "n_atoms": 100000,
        "lr100": float(ob.ref_calc(big, brad, 0, 1.4, 100, 8).sum()),
        "sr1000": float(ob.ref_calc(big, brad, 1, 1.4, 1000, 8).sum()),
    }
    with open(os.path.join(HERE, "totals.json"), "w") as f:
        json.dump(totals, f, indent=1, sort_keys=True)
    print(json.dumps(totals["measured"]["globule100k"]))


if __name__ == "__main__":
    main()

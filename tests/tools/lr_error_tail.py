"""Tail of the fp32 Lee-Richards error on a large structure: the largest per-atom |dSASA| against the fp64 restatement
for several slice counts (the absolute error of a slice scales with its thickness, so low resolutions are the hard case).
Writes gpurun_out/lr_error_tail.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import freesasa_b200 as fs  # noqa: E402
from oracle import bindings as ob  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1_000_000
x, r = fs.workloads.globule(n, seed=5)
if "--pdb-rounded" in sys.argv:  # what a PDB file holds: 3 decimals (exact tangencies between circles become likely)
    x = np.round(x, 3)
    r = np.round(r, 2)
eng = fs.Engine(0)
out = {"atoms": n, "pdb_rounded": "--pdb-rounded" in sys.argv, "cases": {}}
for slices in ((5, 20) if "--quick" in sys.argv else (5, 10, 20, 50, 100)):
    got = eng.calc(fs.LEE_RICHARDS, x, r, 1.4, slices)
    want = ob.oracle_calc(x, r, ob.LEE_RICHARDS, 1.4, slices)
    err = np.abs(got - want)
    top = np.argsort(-err)[:8]
    out["cases"][str(slices)] = {
        "max": float(err.max()), "p999999": float(np.quantile(err, 0.999999)), "n_above_1e-4": int((err > 1e-4).sum()),
        "n_above_5e-4": int((err > 5e-4).sum()), "n_above_1e-3": int((err > 1e-3).sum()),
        "top": [(int(i), float(err[i]), float(want[i]), float(r[i])) for i in top],
    }
    print(slices, json.dumps(out["cases"][str(slices)]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "lr_error_tail%s.json" % ("_pdb" if "--pdb-rounded" in sys.argv else "")), "w"), indent=1)

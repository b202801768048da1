"""First-contact script for the GPU box: prints errors instead of asserting, writes gpurun_out/debug.json."""
import json, os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import freesasa_b200 as fs
from oracle import bindings as ob

out = {}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
try:
    print("available", fs.available(), flush=True)
    f = np.load(os.path.join(ROOT, "tests/golden/synthetic.npz"))
    x, r = f["g3000_xyz"], f["g3000_radii"]
    for prec in (fs.FP32, fs.FP64):
        e = fs.Engine(0, prec)
        start, _ = ob.oracle_neighbours(x, r + 1.4)
        nn = e.neighbour_counts(x, r, 1.4)
        print("prec", prec, "nn mismatch", int((nn != np.diff(start)).sum()), e.stats(), flush=True)
        for alg, res, key in [(0, 20, "lr20"), (0, 100, "lr100"), (1, 100, "sr100"), (1, 1000, "sr1000")]:
            t = time.time(); got = e.calc(alg, x, r, 1.4, res); dt = time.time() - t
            err = np.abs(got - f["g3000_" + key])
            print(f"prec {prec} {key}: max err {err.max():.3e} at {err.argmax()} n_bad(>1e-3) {(err>1e-3).sum()} total {got.sum():.4f} vs {f['g3000_'+key].sum():.4f} wall {dt*1e3:.2f} ms", e.stats(), flush=True)
            out[f"{prec}_{key}"] = float(err.max())
        e.close()
    e = fs.Engine(0)
    x, r = fs.workloads.globule(100000)
    for alg, res in [(0, 100), (1, 1000)]:
        for it in range(3):
            t = time.time(); got = e.calc(alg, x, r, 1.4, res); dt = time.time() - t
            print(f"100k alg {alg}: wall {dt*1e3:.2f} ms", e.stats(), flush=True)
        want = ob.oracle_calc(x, r, alg, 1.4, res)
        err = np.abs(got - want)
        print(f"100k alg {alg}: max err {err.max():.3e} mean {err.mean():.3e} n>1e-4 {(err>1e-4).sum()}", flush=True)
        out[f"100k_{alg}"] = float(err.max())
except Exception:
    traceback.print_exc()
    out["exception"] = traceback.format_exc()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "debug.json"), "w"), indent=1)

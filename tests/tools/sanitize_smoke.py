"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tests/tools/sanitize_smoke.py
Covers both integrators, both precisions, a batch, the crowded (overflow + unstaged) path and sharding."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import freesasa_b200 as fs  # noqa: E402

rng = np.random.default_rng(0)
x, r = fs.workloads.globule(1500, seed=3)
crowd = (rng.uniform(-4, 4, (700, 3)), rng.choice([1.2, 1.9], 700))
for prec in (fs.FP32, fs.FP64):
    e = fs.Engine(0, prec)
    for alg, res in ((0, 10), (1, 64)):
        a = e.calc(alg, x, r, 1.4, res)
        b = e.calc(alg, crowd[0], crowd[1], 1.4, res)
        assert np.isfinite(a).all() and np.isfinite(b).all()
    e.calc_batch(0, fs.workloads.batch(3, 200, 400, seed=1), 1.4, 10)
    e.close()
print("sanitize smoke ok")

"""Measure every BASELINE.json configuration on ONE GPU (C4/C5 are 8-GPU configs; here their
single-GPU time, which bench.py --gpus N scales by sharding).  Writes gpurun_out/configs.json.

    python tests/tools/measure_configs.py [--skip-oracle]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import freesasa_b200 as fs  # noqa: E402
from freesasa_b200 import workloads  # noqa: E402
from oracle import bindings as ob  # noqa: E402


def best_of(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t)
    return min(ts), out


def main():
    skip_oracle = "--skip-oracle" in sys.argv
    out = {}
    eng = fs.Engine(0)
    threads = min(16, os.cpu_count() or 1)
    if ob.ref_available():
        ob.ref_lib().freesasa_set_verbosity(1)

    # C1: one small PDB (3bzd_trimmed, 2754 atoms), LR n_slices=20
    f = np.load(os.path.join(ROOT, "tests/golden/pdb_fixtures.npz"))
    x, r = f["3bzd_trimmed_xyz"], f["3bzd_trimmed_radii"]
    p = fs.Parameters(fs.LEE_RICHARDS, 1.4, 100, 20, 1)
    t, res = best_of(lambda: fs.calc_coord(x, r, p), 20)
    t_ref = None
    if ob.ref_available():
        t0 = time.perf_counter(); ob.ref_calc(x, r, 0, 1.4, 20, 1); t_ref = time.perf_counter() - t0
    out["C1"] = {"atoms": len(r), "e2e_ms": t * 1e3, "atoms_per_s": len(r) / t, "device_ms": eng.stats()["device_ms"],
                 "max_err": float(np.abs(res.sasa - f["3bzd_trimmed_lr20"]).max()), "ref_1thread_ms": None if t_ref is None else t_ref * 1e3}

    # C2 / C3: 100k globule
    x, r = workloads.globule(100000)
    for key, alg, resn in [("C2", 0, 100), ("C3", 1, 1000)]:
        t, got = best_of(lambda: eng.calc(alg, x, r, 1.4, resn), 5)
        st = eng.stats()
        e = {"atoms": len(r), "e2e_ms": t * 1e3, "atoms_per_s": len(r) / t, "device_ms": st["device_ms"], "integrate_ms": st["integrate_ms"]}
        if not skip_oracle:
            t0 = time.perf_counter(); want = ob.ref_calc(x, r, alg, 1.4, resn, threads) if ob.ref_available() else ob.oracle_calc(x, r, alg, 1.4, resn); e["ref_ms"] = (time.perf_counter() - t0) * 1e3
            e["ref_threads"] = threads
            e["max_err"] = float(np.abs(got - want).max())
        out[key] = e

    # C4: 1024 structures of ~5k atoms, LR n_slices=50, one batched pass
    structs = workloads.batch(1024, 4000, 6000, seed=0)
    total = sum(len(rr) for _, rr in structs)
    t, outs = best_of(lambda: eng.calc_batch(0, structs, 1.4, 50), 3)
    st = eng.stats()
    e = {"structures": 1024, "atoms": total, "e2e_ms": t * 1e3, "atoms_per_s": total / t, "device_ms": st["device_ms"], "integrate_ms": st["integrate_ms"]}
    # the reference-facing batch call (freesasa_calc_coord_batch -> fsb200_calc_batch): sub-batches overlapped on two contexts
    t2, res2 = best_of(lambda: fs.calc_batch(0, structs, 1.4, 50), 3)
    e["e2e_overlapped_ms"] = t2 * 1e3
    e["atoms_per_s_overlapped"] = total / t2
    e["overlapped_equals_single_pass"] = bool(all(np.array_equal(res2[k], outs[k]) for k in range(1024)))
    if not skip_oracle:
        errs = [float(np.abs(outs[k] - ob.oracle_calc(structs[k][0], structs[k][1], 0, 1.4, 50)).max()) for k in range(0, 1024, 64)]
        e["max_err_sampled_16_structures"] = max(errs)
    out["C4"] = e

    # C5: one 1M-atom capsid shell, LR n_slices=100
    x, r = workloads.capsid(1_000_000)
    t, got = best_of(lambda: eng.calc(0, x, r, 1.4, 100), 3)
    st = eng.stats()
    e = {"atoms": len(r), "e2e_ms": t * 1e3, "atoms_per_s": len(r) / t, "device_ms": st["device_ms"], "integrate_ms": st["integrate_ms"], "n_items": st["n_items"]}
    if not skip_oracle:
        t0 = time.perf_counter(); want = ob.oracle_calc(x, r, 0, 1.4, 100); e["oracle_ms"] = (time.perf_counter() - t0) * 1e3
        e["max_err"] = float(np.abs(got - want).max())
        e["total"] = float(got.sum()); e["total_oracle"] = float(want.sum())
    out["C5"] = e

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

"""File-to-areas pipeline on the GPU box: PDB text -> structure (row f-1) -> SASA (hot path) -> result tree (row f-3),
this repo's host layer + B200 engine beside the compiled reference on the box's host cores, stage by stage, on the same
bytes.  Writes gpurun_out/pipeline.json.

    python tests/tools/measure_pipeline.py [--quick]
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import freesasa_b200 as fs  # noqa: E402
from freesasa_b200 import structure as st  # noqa: E402
from freesasa_b200 import workloads as w  # noqa: E402
from oracle import bindings as ob  # noqa: E402


def timed(fn, reps, release=None):
    """Best wall time of ``reps`` calls in ms; every result but the last is handed to ``release``."""
    best, out = 1e30, None
    for k in range(reps):
        if out is not None and release:
            release(out)
        t = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t)
    return best * 1e3, out


def stages(api, path, params, reps):
    """ms for read (fopen + freesasa_structure_from_pdb on a file in /dev/shm) / calc / tree / tree free, and the
    per-atom result."""
    st.TreeAPI(api)
    L = api.lib
    L.freesasa_calc_structure.restype = ctypes.POINTER(api.Result)
    t_read, s = timed(lambda: api.from_pdb_path(path), reps, lambda x: x.free())
    t_calc, res = timed(lambda: L.freesasa_calc_structure(s.h, ctypes.byref(params)), reps, L.freesasa_result_free)
    if not res:
        raise RuntimeError("freesasa_calc_structure returned NULL")
    t_tree, root = timed(lambda: L.freesasa_tree_init(res, s.h, b"pipeline"), reps, L.freesasa_node_free)
    t0 = time.perf_counter()
    L.freesasa_node_free(root)
    t_free = (time.perf_counter() - t0) * 1e3
    sasa = np.ctypeslib.as_array(res.contents.sasa, shape=(s.n,)).copy()
    L.freesasa_result_free(res)
    return {"atoms": s.n, "read_ms": t_read, "calc_ms": t_calc, "tree_ms": t_tree, "tree_free_ms": t_free,
            "total_ms": t_read + t_calc + t_tree}, sasa


def ensemble_side(api, path, params, batch, reps, tree_batch=False):
    """read (freesasa_structure_array) / calc / one tree per model, best of `reps`; `batch`: all models in one
    freesasa_calc_structure_batch() call (this repo), else one freesasa_calc_structure() per model (the CLI's loop)."""
    L = api.lib
    st.TreeAPI(api)
    res_p = ctypes.POINTER(api.Result)
    L.freesasa_calc_structure.restype = res_p
    if batch:
        L.freesasa_calc_structure_batch.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(api.Parameters),
                                                    ctypes.POINTER(res_p)]
    best = {"read_ms": 1e30, "calc_ms": 1e30, "tree_ms": 1e30}
    totals = None
    for _ in range(reps):
        t0 = time.perf_counter()
        structures = api.array_path(path, None, st.SEPARATE_MODELS)
        t1 = time.perf_counter()
        n = len(structures)
        results = (res_p * n)()
        if tree_batch:  # calculation and trees in one call (additive freesasa_calc_tree_batch)
            L.freesasa_calc_tree_batch.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(api.Parameters),
                                                   ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_void_p)]
            handles = (ctypes.c_void_p * n)(*[s.h for s in structures])
            tree_arr = (ctypes.c_void_p * n)()
            assert L.freesasa_calc_tree_batch(n, handles, ctypes.byref(params), None, tree_arr) == 0
            t3 = t2 = time.perf_counter()
            totals = [L.freesasa_node_area(L.freesasa_node_children(L.freesasa_node_children(tree_arr[k]))).contents.total for k in range(n)]
            for k in range(n):
                L.freesasa_node_free(tree_arr[k])
            for s in structures:
                s.free()
            best["read_ms"] = min(best["read_ms"], (t1 - t0) * 1e3)
            best["calc_ms"] = min(best["calc_ms"], (t2 - t1) * 1e3)
            best["tree_ms"] = 0.0
            continue
        if batch:
            handles = (ctypes.c_void_p * n)(*[s.h for s in structures])
            assert L.freesasa_calc_structure_batch(n, handles, ctypes.byref(params), results) == 0
        else:
            for k in range(n):
                results[k] = L.freesasa_calc_structure(structures[k].h, ctypes.byref(params))
        t2 = time.perf_counter()
        trees = [L.freesasa_tree_init(results[k], structures[k].h, b"m") for k in range(n)]
        t3 = time.perf_counter()
        totals = [L.freesasa_node_area(L.freesasa_node_children(L.freesasa_node_children(trees[k]))).contents.total for k in range(n)]
        for k in range(n):
            L.freesasa_node_free(trees[k])
            L.freesasa_result_free(results[k])
        for s in structures:
            s.free()
        for key, v in (("read_ms", t1 - t0), ("calc_ms", t2 - t1), ("tree_ms", t3 - t2)):
            best[key] = min(best[key], v * 1e3)
    best["total_ms"] = best["read_ms"] + best["calc_ms"] + best["tree_ms"]
    best["models"] = len(totals)
    return best, totals


def ensemble(mine, ref, threads, n_models=64, n_atoms=5000):
    """An NMR-ensemble-like file (n_models x n_atoms): freesasa_structure_array -> SASA -> one tree per model.
    This repo: all models in ONE device pass; reference: one calculation per structure (src/main.cc:334-362), `threads` threads."""
    text = w.pdb_text(n_atoms, seed=2, chains=2, models=n_models).encode()
    path = "/dev/shm/_fsb_ensemble.pdb"
    with open(path, "wb") as f:
        f.write(text)
    out = {"models": n_models, "bytes": len(text)}
    out["this_repo"], totals = ensemble_side(mine, path, fs.Parameters(fs.LEE_RICHARDS, 1.4, 100, 20, 1), True, 4)
    fused, totals_f = ensemble_side(mine, path, fs.Parameters(fs.LEE_RICHARDS, 1.4, 100, 20, 1), True, 4, tree_batch=True)
    fused["note"] = "calc_ms = freesasa_calc_tree_batch(): device pass + all trees"
    out["this_repo_tree_batch"] = fused
    assert totals_f == totals
    out["reference"], ref_totals = ensemble_side(ref, path, ob.RefParameters(fs.LEE_RICHARDS, 1.4, 100, 20, threads), False, 1)
    out["max_abs_err_total"] = float(max(abs(a - b) for a, b in zip(totals, ref_totals)))
    out["speedup_total"] = out["reference"]["total_ms"] / out["this_repo"]["total_ms"]
    out["speedup_total_tree_batch"] = out["reference"]["total_ms"] / fused["total_ms"]
    os.remove(path)
    return out


def main():
    quick = "--quick" in sys.argv
    threads = min(16, os.cpu_count() or 1)
    mine = st.api()
    ref = st.StructureAPI(ob.ref_lib(), ob.RefResult, ob.RefParameters)
    for api in (mine, ref):
        api.lib.freesasa_set_verbosity(1)
        st.TreeAPI(api)  # declares the ctypes signatures of the node API on the library (pointers are 64-bit)
    out = {"host_threads_reference": threads, "host_cpus": os.cpu_count(), "cases": {}}
    for name, n_atoms, chains in [("14k (2isk-sized)", 13928, 4), ("100k", 100000, 8)] + ([] if quick else [("1M", 1000000, 60)]):
        text = w.pdb_text(n_atoms, seed=5, chains=chains).encode()
        path = "/dev/shm/_fsb_pipeline_%d.pdb" % n_atoms
        with open(path, "wb") as f:
            f.write(text)
        case = {"bytes": len(text)}
        for alg, res_n, key in [(fs.LEE_RICHARDS, 100, "LR-100"), (fs.LEE_RICHARDS, 20, "LR-20")]:
            reps = 3 if n_atoms <= 100000 else 1
            m, sasa_m = stages(mine, path, fs.Parameters(alg, 1.4, res_n, res_n, 1), reps + 2)
            entry = {"this_repo": m}
            if n_atoms <= 100000 or key == "LR-20":
                r, sasa_r = stages(ref, path, ob.RefParameters(alg, 1.4, res_n, res_n, threads), 1 if n_atoms > 20000 else reps)
                entry["reference"] = r
                entry["max_abs_err"] = float(np.abs(sasa_m - sasa_r).max())
                entry["speedup_total"] = r["total_ms"] / m["total_ms"]
            case[key] = entry
            print(name, key, json.dumps(entry), flush=True)
        out["cases"][name] = case
        os.remove(path)
    out["ensemble_64x5k_LR20"] = ensemble(mine, ref, threads)
    print("ensemble", json.dumps(out["ensemble_64x5k_LR20"]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "pipeline.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

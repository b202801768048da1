#!/usr/bin/env python
"""A/B measurement of build-time variants of the integration kernel (knobs in freesasa_b200/csrc/engine.cuh).

    python tests/tools/ab_variants.py --build          # here (no GPU needed): compile every variant in-tree
    python tests/tools/ab_variants.py --run [names]    # on the GPU box: time each variant, write gpurun_out/ab_variants.json

Each variant runs in its own process (FSB200_ENGINE_LIB selects the library).  Reported per variant: kernel time
(CUDA events inside the engine, median of 15 calls on device-resident inputs) for C2 (100k globule, L&R 100), C2 without
the certificate, C3 (S&R 1000), a 64 x 5k batch (L&R 50), and the largest |dSASA| against the fp64 oracle on a 60k-atom
globule with coordinates rounded to PDB precision at n_slices = 5, 20, 100.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

VARIANTS = {
    "product": [],
    "fused": ["env:FSB200_PIPELINE=fused"],            # the same library, everything inside k_integrate (round 1's layout)
    "slices_k3only": ["FSB200_SLICES_K=1"],
    "no_exact_slices": ["FSB200_EXACT_SLICES=0"],
    "r1_kernel": ["@integrate.cu=f4eaa4c"],          # integrate.cu of the round-1 commit, against today's api.cu / cells.cu
}
TAIL_VARIANTS = ("product",)   # these also measure the 1M-atom PDB-rounded error tail


def lib_path(name):
    own = [d for d in VARIANTS[name] if not d.startswith("env:")]
    return os.path.join(ROOT, "freesasa_b200", "csrc", f"libfsb200_{name}.so" if own else "libfsb200.so")


def build():
    from freesasa_b200 import build as b

    b.build_library()
    for name, defs in VARIANTS.items():
        if not [d for d in defs if not d.startswith("env:")]:
            continue
        replace = {}
        for d in [d for d in defs if d.startswith("@")]:
            src, rev = d[1:].split("=")
            old = os.path.join("/tmp", f"{rev}_{src}")
            with open(old, "w") as f:
                f.write(subprocess.run(["git", "show", f"{rev}:freesasa_b200/csrc/{src}"], cwd=ROOT, check=True,
                                       capture_output=True, text=True).stdout)
            replace[src] = old
        print(b.build_variant(name, [d for d in defs if d[0] not in "@e" or not (d.startswith("@") or d.startswith("env:"))], replace=replace), flush=True)


def measure():
    import numpy as np
    import torch

    import freesasa_b200 as fs
    from oracle import bindings as ob

    dev = torch.device("cuda", 0)
    eng = fs.Engine(0)
    out = {}

    def timed(alg, dx, dr, res, offsets=None, reps=15):
        ms = []
        for _ in range(reps + 3):
            eng.calc_device(alg, dx, dr, 1.4, res, offsets=offsets)
            ms.append(eng.stats()["integrate_ms"])
        return float(np.median(ms[3:]))

    x, r = fs.workloads.globule(100000)
    dx, dr = torch.tensor(x, device=dev), torch.tensor(r, device=dev)
    out["c2_lr100_ms"] = timed(0, dx, dr, 100)
    out["c2_certified"] = eng.stats()["n_certified"]
    out["c3_sr1000_ms"] = timed(1, dx, dr, 1000)
    out["c2_lr20_ms"] = timed(0, dx, dr, 20)
    eng.set_certificate(False)
    out["c2_lr100_nocert_ms"] = timed(0, dx, dr, 100, reps=5)
    eng.set_certificate(True)
    structs = fs.workloads.batch(64, 4000, 6000, seed=0)
    bx = torch.tensor(np.concatenate([a for a, _ in structs]), device=dev)
    br = torch.tensor(np.concatenate([b for _, b in structs]), device=dev)
    off = np.concatenate([[0], np.cumsum([len(b) for _, b in structs])]).astype(np.int32)
    out["batch64x5k_lr50_ms"] = timed(0, bx, br, 50, offsets=off)
    xs, rs = fs.workloads.capsid(200000, r_out=120.0, seed=1)
    sx, sr = torch.tensor(xs, device=dev), torch.tensor(rs, device=dev)
    out["shell200k_lr100_ms"] = timed(0, sx, sr, 100, reps=7)
    out["shell200k_certified"] = eng.stats()["n_certified"]
    xe, re_ = fs.workloads.globule(60000, seed=5)
    xe, re_ = np.round(xe, 3), np.round(re_, 2)
    for n_slices in (5, 20, 100):
        got = eng.calc(0, xe, re_, 1.4, n_slices)
        want = ob.oracle_calc(xe, re_, 0, 1.4, n_slices)
        out[f"err_pdb60k_n{n_slices}"] = float(np.abs(got - want).max())
    if os.environ.get("AB_TAIL"):
        xt, rt = fs.workloads.globule(1_000_000, seed=5)
        xt, rt = np.round(xt, 3), np.round(rt, 2)
        for n_slices in (5, 20, 100):
            got = eng.calc(0, xt, rt, 1.4, n_slices)
            out[f"ms_pdb1M_n{n_slices}"] = eng.stats()["integrate_ms"]
            err = np.abs(got - ob.oracle_calc(xt, rt, 0, 1.4, n_slices))
            out[f"err_pdb1M_n{n_slices}"] = [float(err.max()), int((err > 1e-4).sum())]
    print("AB " + json.dumps(out), flush=True)


def run(names):
    results = {}
    for name in names:
        env = dict(os.environ, FSB200_ENGINE_LIB=lib_path(name))
        for d in VARIANTS[name]:
            if d.startswith("env:"):
                k, v = d[4:].split("=", 1)
                env[k] = v
        if name in TAIL_VARIANTS:
            env["AB_TAIL"] = "1"
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--measure"], env=env, capture_output=True, text=True, timeout=900)
        line = [l for l in p.stdout.splitlines() if l.startswith("AB ")]
        results[name] = json.loads(line[-1][3:]) if line else {"failed": (p.stderr or p.stdout)[-2000:]}
        print(name, json.dumps(results[name]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "ab_variants.json"), "w"), indent=1)


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
    elif "--measure" in sys.argv:
        measure()
    else:
        names = [a for a in sys.argv[1:] if not a.startswith("--")] or list(VARIANTS)
        run(names)

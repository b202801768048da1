#!/usr/bin/env python
"""CPU model of the fp32 Lee-Richards slice loop (lr_atom_fastk in freesasa_b200/csrc/integrate.cu): how a slice ends
(buried by one circle / every sector inside a single arc / needs the exact merge / free) and after HOW MANY groups of 32
neighbours that is known, for different orders of the neighbour records.  numpy only (plus the oracle's neighbour list and
the certificate's direction table, a pure host function): design aid, run without a GPU.

    python tests/tools/slice_loop_model.py [n_atoms] [n_slices]

The kernel evaluates all K = 3 groups of a z-sorted list before its votes.  Burial and "every sector covered" are monotone
in the set of arcs, so evaluating the groups one after the other and leaving at the first group that settles the slice is
exact; the model measures how often the first group (the 32 widest caps) is enough.
"""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/tests/", 1)[0])
import freesasa_b200 as fs  # noqa: E402
from oracle import bindings as ob  # noqa: E402


def cert_dirs():
    import ctypes

    L = fs._engine_lib()
    L.fsb200_cert_directions.argtypes = [ctypes.POINTER(ctypes.c_double)]
    out = np.empty(384)
    L.fsb200_cert_directions(out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out.reshape(128, 3)


def certified(D, Ri, Rj, U):
    """certificate of integrate.cu::certify_buried without its fp32 margins (model)."""
    d2 = (D * D).sum(1)
    d = np.sqrt(d2)
    t = (Ri * Ri + d2 - Rj * Rj) / (2 * Ri)
    rho = np.deg2rad(12.5)
    ok = t < d * np.cos(rho)
    if not ok.any():
        return False
    tp = t[ok] * np.cos(rho) + np.sqrt(np.maximum(d2[ok] - t[ok] ** 2, 0)) * np.sin(rho)
    return bool(((U @ D[ok].T) >= tp[None, :]).any(1).all())


def slice_classes(D, Ri, Rj, ns, order):
    """per slice: group index (0-based) after which the slice is settled as buried / fully covered, or -1 if it needs the
    merge (or is free).  order = permutation of the neighbours (group g = order[32g:32g+32])."""
    nn = len(Rj)
    dz, dxy = D[:, 2], np.hypot(D[:, 0], D[:, 1])
    beta = (np.arctan2(D[:, 1], D[:, 0]) + np.pi) * (16 / np.pi)
    delta = 2 * Ri / ns
    zr = -Ri + (np.arange(ns) + 0.5) * delta
    a = np.sqrt(np.maximum((Ri - np.abs(zr)) * (Ri + np.abs(zr)), 0))[:, None]
    dj = np.abs(dz[None, :] - zr[:, None])
    b2 = (Rj[None, :] - dj) * (Rj[None, :] + dj)
    act = b2 > 0
    b = np.sqrt(np.maximum(b2, 0))
    d = dxy[None, :]
    f1, f3, f2 = a + b - d, d + a - b, d + b - a
    touch = act & (f1 > 0)
    bur = touch & (f3 < 0)
    has = touch & ~(f3 < 0) & ~(f2 < 0)
    N, Dn = f1 * f2, f3 * (a + b + d)
    with np.errstate(divide="ignore", invalid="ignore"):
        alpha = np.where(has, 2 * np.arctan2(np.sqrt(np.maximum(N, 0)), np.sqrt(np.maximum(Dn, 0))), 0) * (16 / np.pi)
    st = np.mod(beta[None, :] - alpha, 32)
    en = st + 2 * alpha
    js = np.ceil(st).astype(np.int64)
    cnt = np.floor(en).astype(np.int64) - js
    cnt = np.where(has & (cnt > 0), np.minimum(cnt, 32), 0)
    base = (np.uint64(1) << cnt.astype(np.uint64)) - np.uint64(1)
    sh = (js % 32).astype(np.uint64)
    wide = base << sh
    mask = ((wide | (wide >> np.uint64(32))) & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    G = (nn + 31) // 32
    settled = np.full(ns, -1)
    kind = np.zeros(ns, dtype=np.int64)  # 0 open, 1 buried, 2 covered
    cum = np.zeros(ns, dtype=np.uint64)
    active_groups = np.zeros((ns, G), dtype=bool)
    for g in range(G):
        idx = order[32 * g:32 * g + 32]
        active_groups[:, g] = act[:, idx].any(1)
        gb = bur[:, idx].any(1)
        cum |= np.bitwise_or.reduce(mask[:, idx], axis=1)
        full = cum == np.uint64(0xFFFFFFFF)
        new = (settled < 0) & (gb | full)
        settled[new] = g
        kind[new & gb] = 1
        kind[new & ~gb] = 2
    return settled, kind, active_groups, G


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    ns = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    xyz, radii = fs.workloads.globule(n)
    R = radii + 1.4
    start, lst = ob.oracle_neighbours(xyz, R)
    U = cert_dirs()
    rng = np.random.default_rng(0)
    # candidates for integration: everything the certificate (model) does not settle; sample them
    near_surface = np.where(np.linalg.norm(xyz, axis=1) > np.linalg.norm(xyz, axis=1).max() - 12.0)[0]
    sample = rng.choice(near_surface, size=min(6000, len(near_surface)), replace=False)
    stats = {}
    n_int = 0
    for i in sample:
        nb = lst[start[i]:start[i + 1]]
        if len(nb) == 0:
            continue
        D, Rj, Ri = xyz[nb] - xyz[i], R[nb], R[i]
        if certified(D, Ri, Rj, U) or len(nb) > 96:
            continue
        n_int += 1
        d = np.linalg.norm(D, axis=1)
        t = (Ri * Ri + d * d - Rj * Rj) / (2 * Ri)
        orders = {
            "z (kernel today)": np.argsort(D[:, 2], kind="stable"),
            "cap width t/d": np.argsort(t / d, kind="stable"),
            "t": np.argsort(t, kind="stable"),
            "distance": np.argsort(d, kind="stable"),
        }
        for name, order in orders.items():
            settled, kind, actg, G = slice_classes(D, Ri, Rj, ns, order)
            s = stats.setdefault(name, {"slices": 0, "after": np.zeros(4, dtype=np.int64), "open": 0, "groups_all": 0,
                                        "groups_active": 0, "groups_incremental": 0, "buried": 0, "covered": 0})
            s["slices"] += ns
            for g in range(3):
                s["after"][g] += int((settled == g).sum())
            s["open"] += int((settled < 0).sum())
            s["buried"] += int((kind == 1).sum())
            s["covered"] += int((kind == 2).sum())
            s["groups_all"] += ns * G
            s["groups_active"] += int(actg.sum())
            # incremental: groups evaluated until settled (all G for open slices), counting only z-active groups
            upto = np.where(settled >= 0, settled + 1, G)
            s["groups_incremental"] += int(sum(actg[k, :upto[k]].sum() for k in range(ns)))
    print(f"atoms sampled near the surface: {len(sample)}, not certified (integrated): {n_int}, n_slices {ns}")
    for name, s in stats.items():
        sl = s["slices"]
        print(f"{name:>18}: settled after group 0/1/2: {s['after'][0] / sl:.3f} {s['after'][1] / sl:.3f} {s['after'][2] / sl:.3f}"
              f"  open {s['open'] / sl:.3f}  (buried {s['buried'] / sl:.3f}, covered {s['covered'] / sl:.3f})"
              f"  groups/slice: all {s['groups_all'] / sl:.2f}, z-active {s['groups_active'] / sl:.2f},"
              f" incremental {s['groups_incremental'] / sl:.2f}")


if __name__ == "__main__":
    main()

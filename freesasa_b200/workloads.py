"""Seeded synthetic inputs for the benchmark configurations in BASELINE.json / SURVEY.md §8(d).

numpy only — usable without a GPU and without the CUDA library.  Every generator returns
``xyz`` as float64 ``[n, 3]`` (the reference's AoS coordinate layout, src/coord.h:26-38) and
``radii`` as float64 ``[n]`` (van der Waals radii WITHOUT the probe, as freesasa_calc_coord takes
them, src/freesasa.c:122-142).
"""
from __future__ import annotations

import numpy as np

# ProtOr-like heavy-atom radii (C aliphatic x3, C aromatic/carbonyl, N, N+, O, O-, S) — the ten
# entry table from SURVEY.md §8(d).
RADIUS_TABLE = np.array([1.88, 1.88, 1.88, 1.61, 1.76, 1.64, 1.64, 1.42, 1.46, 1.77])
LATTICE = 2.6  # Å; 17.6 Å^3 per atom ~ protein heavy-atom density
JITTER = 0.6  # Å, uniform +-


def _lattice_in_ball(r_out: float, r_in: float = 0.0):
    m = int(np.ceil(r_out / LATTICE)) + 1
    ax = np.arange(-m, m + 1, dtype=np.float64) * LATTICE
    g = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), axis=-1).reshape(-1, 3)  # scan order i,j,k
    d = np.sqrt((g * g).sum(1))
    keep = (d <= r_out) & (d >= r_in)
    return g[keep], d[keep]


def globule(n: int, seed: int = 0, offset=(0.0, 0.0, 0.0), shuffle: bool = False):
    """Dense globular pseudo-protein of exactly ``n`` atoms: jittered cubic lattice clipped to a
    ball (configs C2/C3 with n = 100000; the unit of C4 with n ~ 5000)."""
    rng = np.random.default_rng(np.random.PCG64(0x9E3779B97F4A7C15 ^ (seed * 0x100000001B3)))
    r = (3.0 * n * LATTICE**3 / (4.0 * np.pi)) ** (1.0 / 3.0)
    pts, d = _lattice_in_ball(r + 2 * LATTICE)
    order = np.argsort(d, kind="stable")[:n]
    order.sort()  # back to lattice scan order
    xyz = pts[order] + rng.uniform(-JITTER, JITTER, size=(n, 3))
    radii = RADIUS_TABLE[rng.integers(0, len(RADIUS_TABLE), size=n)]
    if shuffle:
        p = rng.permutation(n)
        xyz, radii = xyz[p], radii[p]
    xyz = xyz + np.asarray(offset, dtype=np.float64)
    return np.ascontiguousarray(xyz), np.ascontiguousarray(radii)


def capsid(n: int = 1_000_000, r_out: float = 250.0, seed: int = 0):
    """Hollow shell of ``n`` atoms (config C5): the same jittered lattice restricted to
    r_in <= r <= r_out with the shell volume = n * 17.58 Å^3."""
    rng = np.random.default_rng(np.random.PCG64(0xC2B2AE3D27D4EB4F ^ (seed * 0x100000001B3)))
    vol = n * LATTICE**3
    r_in3 = r_out**3 - 3.0 * vol / (4.0 * np.pi)
    r_in = max(r_in3, 0.0) ** (1.0 / 3.0)
    pts, d = _lattice_in_ball(r_out + LATTICE, max(r_in - 2 * LATTICE, 0.0))
    order = np.argsort(-d, kind="stable")[:n]  # outermost n lattice points
    order.sort()
    xyz = pts[order] + rng.uniform(-JITTER, JITTER, size=(len(order), 3))
    radii = RADIUS_TABLE[rng.integers(0, len(RADIUS_TABLE), size=len(order))]
    return np.ascontiguousarray(xyz), np.ascontiguousarray(radii)


def batch(n_struct: int, n_lo: int = 4000, n_hi: int = 6000, seed: int = 0):
    """``n_struct`` independent globules with sizes ~U[n_lo, n_hi] and a random rigid offset each
    (config C4).  Returns a list of (xyz, radii)."""
    rng = np.random.default_rng(np.random.PCG64(0x165667B19E3779F9 ^ (seed * 0x100000001B3)))
    out = []
    for k in range(n_struct):
        nk = int(rng.integers(n_lo, n_hi + 1))
        off = rng.uniform(-300.0, 300.0, size=3)
        out.append(globule(nk, seed=seed * 100003 + k + 1, offset=off))
    return out

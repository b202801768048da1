"""CPU check of the construction behind lr_mark_marginal() (freesasa_b200/csrc/integrate.cu): the slice planes within
rounding distance of a tangency between two slice circles are found in CLOSED FORM from the intersection circle of the
two spheres.  A numpy transcription of the kernel's fp32 formulas must mark a SUPERSET of what the brute-force test
(q = min(|N|,|D|)/max(|N|,|D|) < q_min for every pair of every slice, fp64) marks.  The CUDA routine itself is exercised
by the -m gpu parity tests (1M atoms, PDB-rounded, n = 5 / 20 / 100)."""
import numpy as np
import pytest

from freesasa_b200 import workloads
from oracle import bindings as ob

f = np.float32


def brute_force(D, Ri, Rj, ns, q_min):
    dz, d = D[:, 2], np.hypot(D[:, 0], D[:, 1])
    zs = -Ri + (np.arange(ns) + 0.5) * (2 * Ri / ns)
    a = np.sqrt(np.maximum((Ri - np.abs(zs)) * (Ri + np.abs(zs)), 0))[:, None]
    dj = np.abs(dz[None, :] - zs[:, None])
    b2 = (Rj[None, :] - dj) * (Rj[None, :] + dj)
    b = np.sqrt(np.maximum(b2, 0))
    N = (a + b - d) * (d + b - a)
    Dn = (d + a - b) * (a + b + d)
    lo, hi = np.minimum(np.abs(N), np.abs(Dn)), np.maximum(np.abs(N), np.abs(Dn))
    return ((b2 > 0) & (lo < q_min * hi)).any(1)


def closed_form(D, Ri, Rj, ns, q_min):
    """lr_mark_marginal, line by line, in float32."""
    Ri, dz, Rj, d = f(Ri), D[:, 2].astype(f), Rj.astype(f), np.hypot(D[:, 0], D[:, 1]).astype(f)
    delta, inv_delta = f(2) * Ri / f(ns), f(ns) / (f(2) * Ri)
    zs = -Ri + (np.arange(ns).astype(f) + f(0.5)) * delta
    D3sq = d * d + dz * dz
    inv_D3 = f(1) / np.sqrt(D3sq)
    t = f(0.5) * (D3sq + (Ri - Rj) * (Ri + Rj)) * inv_D3
    rho2 = (Ri - t) * (Ri + t)
    ok = (D3sq > 0) & (rho2 > 0)
    rho = np.sqrt(np.maximum(rho2, f(0)))
    zc, ext = t * dz * inv_D3, rho * d * inv_D3
    slack = f(4e-6) * (f(1) + f(1) / np.maximum(rho, f(1e-3)))
    flags = np.zeros(ns, bool)
    for j in np.where(ok & (f(2) * ext < f(1e-3)))[0]:
        w = f(1e-4) + slack[j]
        flags |= (zs > zc[j] - ext[j] - w - delta) & (zs < zc[j] + ext[j] + w + delta)   # floor / ceil of the kernel: one slice of slack
    for z in (zc - ext, zc + ext):
        a = np.sqrt(np.maximum((Ri - z) * (Ri + z), f(1e-12)))
        zz = z - dz
        b = np.sqrt(np.maximum((Rj - zz) * (Rj + zz), f(0)))
        Nv, Dv = np.abs((a + b - d) * (d + b - a)), np.abs((d + a - b) * (a + b + d))
        zda = f(2) * z * d / a
        slope = np.where(Nv < Dv, np.abs(f(2) * dz - zda), np.abs(-f(2) * dz - zda))
        eps = np.minimum(np.maximum(f(2) * f(q_min) * np.maximum(Nv, Dv) / np.maximum(slope, f(1e-9)), f(2e-6)) + f(2e-6) + slack, delta)
        sc = np.rint((z + Ri) * inv_delta - f(0.5)).astype(int)
        for k in (-1, 0, 1):
            s = sc + k
            s2 = np.clip(s, 0, ns - 1)
            hit = ok & (s >= 0) & (s < ns) & (np.abs(zs[s2] - z) < eps)
            flags[s2[hit]] = True
    return flags


@pytest.mark.parametrize("ns,q_min", [(100, 3e-6), (20, 3e-6), (5, 2.4e-5)])
def test_closed_form_marks_a_superset_of_the_brute_force_test(ns, q_min):
    xyz, radii = workloads.globule(30000, seed=2)
    xyz, radii = np.round(xyz, 3), np.round(radii, 2)          # PDB precision: exact tangencies become likely
    R = radii + 1.4
    start, lst = ob.oracle_neighbours(xyz, R)
    rad = np.linalg.norm(xyz, axis=1)
    surface = np.where(rad > rad.max() - 8)[0]
    sample = np.random.default_rng(ns).choice(surface, 700, replace=False)
    n_brute = n_closed = 0
    for i in sample:
        nb = lst[start[i]:start[i + 1]]
        if len(nb) == 0:
            continue
        D, Rj = xyz[nb] - xyz[i], R[nb]
        brute, closed = brute_force(D, R[i], Rj, ns, q_min), closed_form(D, R[i], Rj, ns, q_min)
        assert not (brute & ~closed).any(), (int(i), np.where(brute & ~closed)[0])
        n_brute += int(brute.sum())
        n_closed += int(closed.sum())
    assert n_brute > 0                       # the sample does contain marginal slices ...
    assert n_closed < 6 * n_brute + 20       # ... and the closed form does not mark wildly more than needed

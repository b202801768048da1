"""CPU tests: pin the oracle (oracle/oracle.c) to every known answer the reference's own tests
hold for the hot path, to committed fixtures generated from the unmodified reference, and — when
oracle/_ref is built — bit-for-bit to the reference itself."""
import math

import numpy as np
import pytest

from oracle import bindings as ob
from tests import analytic

TWOPI = 2 * math.pi
PDBS = ["1ubq", "2jo4", "3bkr", "5dx9", "3bzd_trimmed", "1d3z"]


# ---- reference tests/test_freesasa.c:155-178,302-332,432-473 : golden totals -------------------
@pytest.mark.parametrize(
    "name,key,alg,res,gold",
    [
        ("1ubq", "lr20", ob.LEE_RICHARDS, 20, 4804.055641),
        ("1ubq", "sr100", ob.SHRAKE_RUPLEY, 100, 4834.716265),
        ("3bzd_trimmed", "sr100", ob.SHRAKE_RUPLEY, 100, 16133.867124),
        ("1d3z", "sr100", ob.SHRAKE_RUPLEY, 100, 5000.340175),
    ],
)
def test_published_totals(pdb_fixtures, name, key, alg, res, gold):
    sasa = ob.oracle_calc(pdb_fixtures[name + "_xyz"], pdb_fixtures[name + "_radii"], alg, 1.4, res)
    assert abs(sasa.sum() - gold) < 1e-5  # the reference's own tolerance
    assert abs(pdb_fixtures[f"{name}_{key}"].sum() - gold) < 1e-5  # fixture really is the reference


# ---- committed per-atom vectors from the unmodified reference ---------------------------------
@pytest.mark.parametrize("name", PDBS)
@pytest.mark.parametrize("key,alg,res", [("lr20", ob.LEE_RICHARDS, 20), ("sr100", ob.SHRAKE_RUPLEY, 100)])
def test_per_atom_pdb(pdb_fixtures, name, key, alg, res):
    sasa = ob.oracle_calc(pdb_fixtures[name + "_xyz"], pdb_fixtures[name + "_radii"], alg, 1.4, res)
    np.testing.assert_array_equal(sasa, pdb_fixtures[f"{name}_{key}"])


@pytest.mark.parametrize(
    "key,alg,res",
    [("lr20", 0, 20), ("lr100", 0, 100), ("sr100", 1, 100), ("sr1000", 1, 1000)],
)
def test_per_atom_synthetic(synthetic_fixtures, key, alg, res):
    f = synthetic_fixtures
    sasa = ob.oracle_calc(f["g3000_xyz"], f["g3000_radii"], alg, 1.4, res)
    np.testing.assert_array_equal(sasa, f["g3000_" + key])


def test_per_atom_far_from_origin(synthetic_fixtures):
    f = synthetic_fixtures
    np.testing.assert_array_equal(ob.oracle_calc(f["off1500_xyz"], f["off1500_radii"], 0, 1.4, 20), f["off1500_lr20"])
    np.testing.assert_array_equal(ob.oracle_calc(f["off1500_xyz"], f["off1500_radii"], 1, 1.4, 100), f["off1500_sr100"])


# ---- reference src/sasa_lr.c:436-475 : arc merge known answers ---------------------------------
@pytest.mark.parametrize(
    "arcs,expect",
    [
        ([0, 0.1 * TWOPI, 0.9 * TWOPI, TWOPI], 0.8 * TWOPI),
        ([0.9 * TWOPI, TWOPI, 0, 0.1 * TWOPI], 0.8 * TWOPI),
        ([0, TWOPI, 1, 2], 0.0),
        ([1, 2, 0, TWOPI], 0.0),
        ([0.1 * TWOPI, 0.2 * TWOPI, 0.5 * TWOPI, 0.6 * TWOPI], 0.8 * TWOPI),
        ([0.1 * TWOPI, 0.3 * TWOPI, 0.15 * TWOPI, 0.2 * TWOPI], 0.8 * TWOPI),
        ([0.15 * TWOPI, 0.2 * TWOPI, 0.1 * TWOPI, 0.3 * TWOPI], 0.8 * TWOPI),
        ([0.05, 0.1, 0.5, 0.6, 0, 0.15, 0.7, 0.8, 0.75, TWOPI], 0.45),
        ([], TWOPI),
    ],
)
def test_exposed_arc_kat(arcs, expect):
    assert abs(ob.oracle_exposed_arc(arcs) - expect) < 1e-10


# ---- reference tests/test_nb.c:7-27 : contact known answers ------------------------------------
def test_contact_kat():
    v = np.array([0, 0, 0, 1, 1, 1, -1, 1, -1, 2, 0, -2, 2, 2, 0, -5, 5, 5], dtype=float)
    r = np.array([4, 2, 2, 2, 2, 2], dtype=float)
    start, lst = ob.oracle_neighbours(v, r)
    row = lambda i: set(lst[start[i] : start[i + 1]].tolist())
    assert 1 in row(0) and 0 in row(1)
    assert 5 not in row(0)
    # symmetric, no self, no duplicates
    for i in range(6):
        assert i not in row(i)
        assert len(row(i)) == start[i + 1] - start[i]
        for j in row(i):
            assert i in row(j)


def test_neighbours_match_brute_force():
    rng = np.random.default_rng(3)
    xyz = rng.uniform(-12, 12, size=(400, 3))
    R = rng.uniform(1.0, 3.5, size=400)
    start, lst = ob.oracle_neighbours(xyz, R)
    d2 = ((xyz[:, None, :] - xyz[None, :, :]) ** 2).sum(-1)
    cut = (R[:, None] + R[None, :]) ** 2
    want = (d2 < cut) & ~np.eye(400, dtype=bool)
    got = np.zeros_like(want)
    for i in range(400):
        got[i, lst[start[i] : start[i + 1]]] = True
    assert (got == want).all()


# ---- reference tests/test_freesasa.c:27-43,59-136 : analytic two spheres and invariances --------
@pytest.mark.parametrize("x1,x2", analytic.TWO_SPHERE_CASES)
def test_two_spheres_analytic(x1, x2):
    xyz, r = np.array([x1, x2], dtype=float), np.array([1.0, 2.0])
    exact = analytic.surface_two_spheres(x1, x2, 1.0, 2.0, 1.4)
    lr = ob.oracle_calc(xyz, r, ob.LEE_RICHARDS, 1.4, 20000).sum()
    sr = ob.oracle_calc(xyz, r, ob.SHRAKE_RUPLEY, 1.4, 5000).sum()
    assert analytic.rel_err(exact, lr) < 1e-5
    assert analytic.rel_err(exact, sr) < 1e-3


@pytest.mark.parametrize("alg,res,tol", [(0, 20000, 1e-5), (1, 5000, 1e-3)])
def test_four_spheres_invariance(alg, res, tol):
    r = np.array(analytic.FOUR_SPHERE_RADII)
    ref = ob.oracle_calc(np.array(analytic.FOUR_SPHERE_POSES[0], dtype=float), r, alg, 1.4, res).sum()
    for pose in analytic.FOUR_SPHERE_POSES[1:]:
        got = ob.oracle_calc(np.array(pose, dtype=float), r, alg, 1.4, res).sum()
        assert analytic.rel_err(ref, got) < tol


def test_single_atom():  # reference tests/test_freesasa.c:138-153
    for alg, res in [(0, 20), (1, 100)]:
        s = ob.oracle_calc(np.zeros((1, 3)), np.array([1.0]), alg, 1.4, res)
        assert abs(s[0] - 4 * math.pi * 2.4 * 2.4) < 1e-9


def test_test_points_on_unit_sphere():
    p = ob.oracle_test_points(1000)
    assert np.allclose((p * p).sum(1), 1.0, atol=1e-12)
    assert abs(p[0, 2] - (1 - 1.0 / 1000)) < 1e-15


# ---- bit-for-bit against the reference itself, seeded random inputs ------------------------------
@pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("alg,res", [(0, 37), (1, 213)])
def test_bit_exact_vs_reference(seed, alg, res):
    rng = np.random.default_rng(seed)
    n = 1200
    xyz = rng.uniform(-18, 18, size=(n, 3)) + rng.uniform(-500, 500, size=3)
    radii = rng.choice([1.2, 1.42, 1.64, 1.88, 2.5, 0.0], size=n)
    # The reference's S&R loop reads nb[i][0] even when atom i has no neighbours (uninitialised
    # heap, src/sasa_sr.c:310-316 with src/nb.c:310) and can crash on isolated atoms; keep only
    # atoms with at least one neighbour so the reference itself is well defined.
    start, _ = ob.oracle_neighbours(xyz, radii + 1.4)
    for _ in range(4):
        keep = np.diff(start) > 0
        xyz, radii = xyz[keep], radii[keep]
        start, _ = ob.oracle_neighbours(xyz, radii + 1.4)
    assert (np.diff(start) > 0).all() and len(radii) > 800
    ob.ref_lib().freesasa_set_verbosity(2)
    for threads in (1, 3):
        ref = ob.ref_calc(xyz, radii, alg, 1.4, res, threads)
        np.testing.assert_array_equal(ob.oracle_calc(xyz, radii, alg, 1.4, res), ref)

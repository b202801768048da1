"""GPU parity tests (run on the B200 box with -m gpu): the CUDA engine, called through the C ABI /
the reference-named C host layer, against the CPU oracle, the committed reference vectors and the
reference's own known answers.

Tolerances (north star): Lee-Richards per-atom |dSASA| <= 1e-3 Å^2 in the default fp32 mode (we
assert 5e-4), <= 1e-8 in fp64 mode; Shrake-Rupley is a point count and must be EXACT (<= 1e-9 Å^2
from the shared fp64 final multiply)."""
import math

import numpy as np
import pytest

import freesasa_b200 as fs
from oracle import bindings as ob
from tests import analytic

pytestmark = pytest.mark.gpu

LR_TOL_FP32 = 5e-4
LR_TOL_FP64 = 1e-8
SR_TOL = 1e-9
PDBS = ["1ubq", "2jo4", "3bkr", "5dx9", "3bzd_trimmed", "1d3z"]


@pytest.fixture(scope="module")
def eng32():
    e = fs.Engine(0, fs.FP32)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng64():
    e = fs.Engine(0, fs.FP64)
    yield e
    e.close()


def params(alg, res, threads=1, probe=1.4):
    return fs.Parameters(alg, probe, res, res, threads)


def maxerr(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max())


# ---- reference golden totals through the reference-named entry point (tests/test_freesasa.c:155-178 ...) ----
@pytest.mark.parametrize(
    "name,alg,res,gold,tol",
    [
        ("1ubq", fs.LEE_RICHARDS, 20, 4804.055641, 2e-3),   # sum of 602 fp32-path areas
        ("1ubq", fs.SHRAKE_RUPLEY, 100, 4834.716265, 1e-5),  # the reference's own tolerance
        ("3bzd_trimmed", fs.SHRAKE_RUPLEY, 100, 16133.867124, 1e-5),
        ("1d3z", fs.SHRAKE_RUPLEY, 100, 5000.340175, 1e-5),
    ],
)
def test_published_totals(pdb_fixtures, name, alg, res, gold, tol):
    r = fs.calc_coord(pdb_fixtures[name + "_xyz"], pdb_fixtures[name + "_radii"], params(alg, res))
    assert abs(r.total - gold) < tol
    assert r.n_atoms == len(pdb_fixtures[name + "_radii"])
    assert abs(r.total - r.sasa.sum()) < 1e-9
    assert r.parameters.alg == alg


def test_thread_count_is_irrelevant(pdb_fixtures):  # tests/test_freesasa.c:404-429
    x, r = pdb_fixtures["1ubq_xyz"], pdb_fixtures["1ubq_radii"]
    a = fs.calc_coord(x, r, params(fs.SHRAKE_RUPLEY, 100, threads=1))
    b = fs.calc_coord(x, r, params(fs.SHRAKE_RUPLEY, 100, threads=2))
    np.testing.assert_array_equal(a.sasa, b.sasa)


# ---- per-atom against committed vectors of the unmodified reference --------------------------------------
@pytest.mark.parametrize("name", PDBS)
def test_per_atom_pdb(pdb_fixtures, eng32, eng64, name):
    x, r = pdb_fixtures[name + "_xyz"], pdb_fixtures[name + "_radii"]
    assert maxerr(eng32.calc(fs.LEE_RICHARDS, x, r, 1.4, 20), pdb_fixtures[name + "_lr20"]) < LR_TOL_FP32
    assert maxerr(eng64.calc(fs.LEE_RICHARDS, x, r, 1.4, 20), pdb_fixtures[name + "_lr20"]) < LR_TOL_FP64
    assert maxerr(eng32.calc(fs.SHRAKE_RUPLEY, x, r, 1.4, 100), pdb_fixtures[name + "_sr100"]) < SR_TOL
    assert maxerr(eng64.calc(fs.SHRAKE_RUPLEY, x, r, 1.4, 100), pdb_fixtures[name + "_sr100"]) < SR_TOL


@pytest.mark.parametrize("key,alg,res", [("lr20", 0, 20), ("lr100", 0, 100), ("sr100", 1, 100), ("sr1000", 1, 1000)])
def test_per_atom_synthetic(synthetic_fixtures, eng32, eng64, key, alg, res):
    f = synthetic_fixtures
    tol32, tol64 = (LR_TOL_FP32, LR_TOL_FP64) if alg == 0 else (SR_TOL, SR_TOL)
    assert maxerr(eng32.calc(alg, f["g3000_xyz"], f["g3000_radii"], 1.4, res), f["g3000_" + key]) < tol32
    assert maxerr(eng64.calc(alg, f["g3000_xyz"], f["g3000_radii"], 1.4, res), f["g3000_" + key]) < tol64


def test_far_from_origin(synthetic_fixtures, eng32):
    """PDB coordinates sit hundreds of Å from the origin; the local-frame fp32 path must not care."""
    f = synthetic_fixtures
    assert maxerr(eng32.calc(0, f["off1500_xyz"], f["off1500_radii"], 1.4, 20), f["off1500_lr20"]) < LR_TOL_FP32
    assert maxerr(eng32.calc(1, f["off1500_xyz"], f["off1500_radii"], 1.4, 100), f["off1500_sr100"]) < SR_TOL
    x = f["off1500_xyz"] + np.array([1.0e5, -2.0e5, 3.0e5])  # exactly representable shift
    assert maxerr(eng32.calc(1, x, f["off1500_radii"], 1.4, 100), ob.oracle_calc(x, f["off1500_radii"], 1, 1.4, 100)) < SR_TOL


# ---- neighbour search row (src/nb.c) ---------------------------------------------------------------------
def test_neighbour_counts(eng32, synthetic_fixtures):
    x, r = synthetic_fixtures["g3000_xyz"], synthetic_fixtures["g3000_radii"]
    start, _ = ob.oracle_neighbours(x, r + 1.4)
    np.testing.assert_array_equal(eng32.neighbour_counts(x, r, 1.4), np.diff(start))
    v = np.array([0, 0, 0, 1, 1, 1, -1, 1, -1, 2, 0, -2, 2, 2, 0, -5, 5, 5], dtype=float)  # tests/test_nb.c:7-27
    rr = np.array([4, 2, 2, 2, 2, 2], dtype=float)
    start, _ = ob.oracle_neighbours(v, rr)
    np.testing.assert_array_equal(eng32.neighbour_counts(v, rr, 0.0), np.diff(start))


# ---- analytic known answers (tests/test_freesasa.c:27-43,59-136) -----------------------------------------
@pytest.mark.parametrize("x1,x2", analytic.TWO_SPHERE_CASES)
def test_two_spheres_analytic(x1, x2):
    xyz, r = np.array([x1, x2], dtype=float), np.array([1.0, 2.0])
    exact = analytic.surface_two_spheres(x1, x2, 1.0, 2.0, 1.4)
    lr = fs.calc_coord(xyz, r, params(fs.LEE_RICHARDS, 20000)).total
    sr = fs.calc_coord(xyz, r, params(fs.SHRAKE_RUPLEY, 5000)).total
    assert analytic.rel_err(exact, lr) < 1e-5
    assert analytic.rel_err(exact, sr) < 1e-3
    assert abs(sr - ob.oracle_calc(xyz, r, 1, 1.4, 5000).sum()) < SR_TOL


@pytest.mark.parametrize("alg,res,tol", [(0, 20000, 1e-5), (1, 5000, 1e-3)])
def test_four_spheres_invariance(alg, res, tol):
    r = np.array(analytic.FOUR_SPHERE_RADII)
    ref = fs.calc_coord(np.array(analytic.FOUR_SPHERE_POSES[0], dtype=float), r, params(alg, res)).total
    for pose in analytic.FOUR_SPHERE_POSES[1:]:
        got = fs.calc_coord(np.array(pose, dtype=float), r, params(alg, res)).total
        assert analytic.rel_err(ref, got) < tol


def test_single_atom():  # tests/test_freesasa.c:138-153
    for alg, res in [(0, 20), (1, 100)]:
        r = fs.calc_coord(np.zeros((1, 3)), np.array([1.0]), params(alg, res))
        assert abs(r.sasa[0] - r.total) < 1e-10
        assert abs(r.total - 4 * math.pi * 2.4 * 2.4) < 1e-4


# ---- edge cases -------------------------------------------------------------------------------------------
def test_edge_cases(eng32, eng64):
    rng = np.random.default_rng(5)
    cases = {
        "two_isolated": (np.array([[0.0, 0, 0], [100.0, 0, 0]]), np.array([1.5, 2.0])),
        "zero_radius_atoms": (rng.uniform(-8, 8, (300, 3)), rng.choice([0.0, 1.5, 1.9], 300)),
        "buried_small_in_big": (np.array([[0.0, 0, 0], [0.3, 0.1, -0.2]]), np.array([5.0, 0.5])),
        "line_along_z": (np.stack([np.zeros(40), np.zeros(40), np.arange(40) * 1.1], 1), np.full(40, 1.7)),
        "flat_sheet": (np.concatenate([rng.uniform(-20, 20, (500, 2)), np.zeros((500, 1))], 1), np.full(500, 1.8)),
        "wide_radii": (rng.uniform(-15, 15, (400, 3)), rng.uniform(0.5, 6.0, 400)),
        "sparse_far_apart": (rng.uniform(-5000, 5000, (64, 3)), np.full(64, 1.8)),
    }
    for name, (x, r) in cases.items():
        for alg, res, tol in [(0, 25, LR_TOL_FP32), (1, 173, SR_TOL)]:
            want = ob.oracle_calc(x, r, alg, 1.4, res)
            assert maxerr(eng32.calc(alg, x, r, 1.4, res), want) < tol, (name, alg)
        assert maxerr(eng64.calc(0, x, r, 1.4, 25), ob.oracle_calc(x, r, 0, 1.4, 25)) < LR_TOL_FP64, name


@pytest.mark.parametrize("res", [1, 2, 3, 7, 31, 32, 33, 257])
def test_odd_resolutions(eng32, eng64, res):
    """Resolutions that are not multiples of the warp width, down to a single slice / test point."""
    x, r = fs.workloads.globule(600, seed=8)
    assert maxerr(eng32.calc(0, x, r, 1.4, res), ob.oracle_calc(x, r, 0, 1.4, res)) < 2e-3 / max(1, min(res, 4)) + LR_TOL_FP32
    assert maxerr(eng64.calc(0, x, r, 1.4, res), ob.oracle_calc(x, r, 0, 1.4, res)) < LR_TOL_FP64
    assert maxerr(eng32.calc(1, x, r, 1.4, res), ob.oracle_calc(x, r, 1, 1.4, res)) < SR_TOL


def test_ragged_batch_with_tiny_structures(eng32):
    """A batch mixing 1-atom, 2-atom and ordinary structures (ragged sizes, cells with a single atom)."""
    rng = np.random.default_rng(12)
    structs = [(np.zeros((1, 3)), np.array([1.7])),
               (np.array([[0.0, 0, 0], [1.0, 0.5, 0.2]]) + 300.0, np.array([1.5, 1.9])),
               fs.workloads.globule(257, seed=1), (rng.uniform(-3, 3, (33, 3)), rng.uniform(1, 2, 33)),
               fs.workloads.globule(1000, seed=2, offset=(-400.0, 250.0, 90.0))]
    for alg, res, tol in [(0, 20, LR_TOL_FP32), (1, 100, SR_TOL)]:
        outs = eng32.calc_batch(alg, structs, 1.4, res)
        for (x, r), got in zip(structs, outs):
            assert got.shape == r.shape
            assert maxerr(got, ob.oracle_calc(x, r, alg, 1.4, res)) < tol


def test_dense_packing_uses_the_wide_paths(eng32, eng64):
    """Twice the protein density (as with explicit hydrogens): 97..160 neighbours take the keyed all-arcs path,
    more than 160 the global-memory overflow kernel; both must agree with the oracle like the common path."""
    x, r = fs.workloads.globule(4000, seed=13)
    x = x * 0.78
    start, _ = ob.oracle_neighbours(x, r + 1.4)
    nn = np.diff(start)
    assert (nn > 96).mean() > 0.3 and nn.max() > 120
    for alg, res, tol in [(0, 24, LR_TOL_FP32), (1, 150, SR_TOL)]:
        assert maxerr(eng32.calc(alg, x, r, 1.4, res), ob.oracle_calc(x, r, alg, 1.4, res)) < tol
    assert maxerr(eng64.calc(0, x, r, 1.4, 24), ob.oracle_calc(x, r, 0, 1.4, 24)) < LR_TOL_FP64
    np.testing.assert_array_equal(eng32.neighbour_counts(x, r, 1.4), nn)


def test_many_tiny_work_items_stress(eng32):
    """Hundreds of small and sparse structures in one pass: most work items hold one or two atoms, so the
    kernel's tile ring turns over thousands of times per CTA.  (Regression: a warp could claim a non-existent
    atom of an exhausted one-atom item whose slot was being recycled; results then varied from run to run.)"""
    rng = np.random.default_rng(21)
    structs = []
    for k in range(400):
        n = int(rng.integers(20, 260))
        if k % 3 == 0:   # sparse: atoms mostly alone in their cells
            structs.append((rng.uniform(-40, 40, (n, 3)), rng.choice([1.5, 1.8], n)))
        else:
            structs.append(fs.workloads.globule(n, seed=1000 + k, offset=rng.uniform(-200, 200, 3)))
    for alg, res, tol in [(0, 12, LR_TOL_FP32), (1, 60, SR_TOL)]:
        first = eng32.calc_batch(alg, structs, 1.4, res)
        for rep in range(3):
            again = eng32.calc_batch(alg, structs, 1.4, res)
            for a, b in zip(first, again):
                np.testing.assert_array_equal(a, b)
        for (x, r), got in zip(structs, first):
            assert maxerr(got, ob.oracle_calc(x, r, alg, 1.4, res)) < tol


def test_concurrent_callers(synthetic_fixtures):
    """The reference library is re-entrant (doc/doxy-main.md:741-756); so must the drop-in be: several host threads
    call freesasa_calc_coord at once (ctypes releases the GIL), each gets a context from the pool."""
    import threading

    f = synthetic_fixtures
    x, r = f["g3000_xyz"], f["g3000_radii"]
    want_lr, want_sr = f["g3000_lr20"], f["g3000_sr100"]   # (NpzFile is not thread-safe: load before the threads start)
    big_x, big_r = fs.workloads.globule(300000, seed=5)   # large enough to wake the host copy pool
    want_big = None
    errors = []

    def worker(k):
        try:
            for rep in range(6):
                if k % 2 == 0:
                    got = fs.calc_coord(x, r, params(fs.LEE_RICHARDS, 20)).sasa
                    assert maxerr(got, want_lr) < LR_TOL_FP32
                else:
                    got = fs.calc_coord(x, r, params(fs.SHRAKE_RUPLEY, 100)).sasa
                    assert maxerr(got, want_sr) < SR_TOL
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    def big_worker():
        try:
            a = fs.calc_coord(big_x, big_r, params(fs.SHRAKE_RUPLEY, 60)).sasa
            b = fs.calc_coord(big_x, big_r, params(fs.SHRAKE_RUPLEY, 60)).sasa
            np.testing.assert_array_equal(a, b)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(6)] + [threading.Thread(target=big_worker) for _ in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_large_probe_and_radii(eng32):
    """Probe and radii far outside the protein range: many more neighbours per atom, coarse grid."""
    x, r = fs.workloads.globule(900, seed=5)
    for probe in (0.0, 3.0, 6.0):
        for alg, res, tol in [(0, 16, LR_TOL_FP32), (1, 100, SR_TOL)]:
            assert maxerr(eng32.calc(alg, x, r, probe, res), ob.oracle_calc(x, r, alg, probe, res)) < tol, (probe, alg)


def test_zero_probe(eng32):
    x, r = fs.workloads.globule(800, seed=3)
    for alg, res, tol in [(0, 20, LR_TOL_FP32), (1, 100, SR_TOL)]:
        assert maxerr(eng32.calc(alg, x, r, 0.0, res), ob.oracle_calc(x, r, alg, 0.0, res)) < tol


def test_crowded_neighbourhoods_take_the_overflow_path(eng32, eng64):
    """> 160 neighbours per atom (smem list capacity) and > 640 atoms per 27-cell tile."""
    rng = np.random.default_rng(9)
    x = rng.uniform(-4.0, 4.0, (1500, 3))
    r = rng.choice([1.2, 1.6, 1.9], 1500)
    for alg, res, tol in [(0, 12, LR_TOL_FP32), (1, 96, SR_TOL)]:
        got = eng32.calc(alg, x, r, 1.4, res)
        st = eng32.stats()
        assert st["n_overflow"] > 0 and st["max_neighbours"] > 160
        assert maxerr(got, ob.oracle_calc(x, r, alg, 1.4, res)) < tol
    start, _ = ob.oracle_neighbours(x, r + 1.4)
    np.testing.assert_array_equal(eng32.neighbour_counts(x, r, 1.4), np.diff(start))
    assert maxerr(eng64.calc(0, x, r, 1.4, 12), ob.oracle_calc(x, r, 0, 1.4, 12)) < LR_TOL_FP64


def test_duplicate_positions_do_not_poison(eng32):
    """Coincident atoms make the reference's acos argument 0/0 (NaN area, src/sasa_lr.c:335); the engine
    treats the coincident circle as a zero-length arc and must stay finite; S&R stays exact."""
    x, r = fs.workloads.globule(400, seed=2)
    x = np.concatenate([x, x[:5]])
    r = np.concatenate([r, r[:5]])
    lr = eng32.calc(0, x, r, 1.4, 20)
    assert np.isfinite(lr).all()
    assert maxerr(eng32.calc(1, x, r, 1.4, 100), ob.oracle_calc(x, r, 1, 1.4, 100)) < SR_TOL


def test_non_finite_input_fails_loudly(eng32):
    x, r = fs.workloads.globule(100, seed=1)
    x[17, 1] = np.nan
    with pytest.raises(RuntimeError, match="non-finite"):
        eng32.calc(0, x, r, 1.4, 20)
    x[17, 1] = 0.0
    assert np.isfinite(eng32.calc(0, x, r, 1.4, 20)).all()  # context still usable


def test_run_to_run_bit_identical(eng32):
    x, r = fs.workloads.globule(5000, seed=4, shuffle=True)
    # call 1: plain launches, call 2: captures the cell-list build into a CUDA graph, calls 3+: replay it
    runs = [eng32.calc(0, x, r, 1.4, 50) for _ in range(4)]
    for other in runs[1:]:
        np.testing.assert_array_equal(runs[0], other)
    assert maxerr(runs[0], ob.oracle_calc(x, r, 0, 1.4, 50)) < LR_TOL_FP32
    c, d = eng32.calc(1, x, r, 1.4, 200), eng32.calc(1, x, r, 1.4, 200)
    np.testing.assert_array_equal(c, d)
    x2 = x + 0.25  # same shape, new coordinates: the replayed graph must read the new upload
    assert maxerr(eng32.calc(1, x2, r, 1.4, 200), ob.oracle_calc(x2, r, 1, 1.4, 200)) < SR_TOL


def test_shuffled_input_order(eng32):
    x, r = fs.workloads.globule(4000, seed=6)
    p = np.random.default_rng(0).permutation(4000)
    a = eng32.calc(1, x, r, 1.4, 100)
    b = eng32.calc(1, x[p], r[p], 1.4, 100)
    np.testing.assert_array_equal(a[p], b)


# ---- batches of independent structures (config C4 shape) -------------------------------------------------
def test_batch_equals_individual(eng32):
    structs = fs.workloads.batch(12, 300, 900, seed=1)
    for alg, res, tol in [(0, 50, LR_TOL_FP32), (1, 100, SR_TOL)]:
        outs = eng32.calc_batch(alg, structs, 1.4, res)
        for (x, r), got in zip(structs, outs):
            assert maxerr(got, ob.oracle_calc(x, r, alg, 1.4, res)) < tol
            np.testing.assert_array_equal(got, eng32.calc(alg, x, r, 1.4, res))
    res = fs.calc_coord_batch(structs, params(fs.LEE_RICHARDS, 50))
    assert len(res) == 12 and all(abs(a.total - a.sasa.sum()) < 1e-9 for a in res)


# ---- device-resident path + sharding of one structure (config C5 shape) -----------------------------------
def test_device_path_and_sharding(eng32):
    import torch

    x, r = fs.workloads.capsid(30000, r_out=60.0, seed=1)
    dx = torch.tensor(x, device="cuda:0")
    dr = torch.tensor(r, device="cuda:0")
    whole = eng32.calc_device(0, dx, dr, 1.4, 30).cpu().numpy()
    assert maxerr(whole, ob.oracle_calc(x, r, 0, 1.4, 30)) < LR_TOL_FP32
    n, k = len(r), 4
    gathered = torch.zeros(n, dtype=torch.float64, device="cuda:0")
    for i in range(k):
        part = eng32.calc_device(0, dx, dr, 1.4, 30, shard=(i, k))
        b, e = eng32.shard_range(n, i, k)
        gathered[b:e] = part[b:e]
    out = eng32.unpermute(gathered).cpu().numpy()
    np.testing.assert_array_equal(out, whole)


# ---- full-size benchmark configurations (C2, C3) ------------------------------------------------------------
def test_full_size_100k(eng32, totals):
    x, r = fs.workloads.globule(100000)
    lr = eng32.calc(0, x, r, 1.4, 100)
    want = ob.oracle_calc(x, r, 0, 1.4, 100)
    assert maxerr(lr, want) < 1e-3  # the north-star tolerance at the headline configuration
    assert abs(want.sum() - totals["measured"]["globule100k"]["lr100"]) < 1e-6
    sr = eng32.calc(1, x, r, 1.4, 1000)
    want_sr = ob.oracle_calc(x, r, 1, 1.4, 1000)
    assert maxerr(sr, want_sr) < SR_TOL
    assert abs(want_sr.sum() - totals["measured"]["globule100k"]["sr1000"]) < 1e-6
    # size-independent properties: isolated far copy changes nothing; total bounded by sum of spheres
    assert (lr >= 0).all() and (lr <= 4 * math.pi * (r + 1.4) ** 2 + 1e-6).all()


def test_large_batch_overlapped_on_two_contexts_equals_one_pass(eng32):
    """fsb200_calc_batch() cuts batches of >= 400k atoms into sub-batches worked through by two contexts from two
    threads (transfers hidden behind kernels): results must not depend on the split, and errors must surface."""
    structs = fs.workloads.batch(120, 3000, 5000, seed=5)          # ~480k atoms -> overlapped path
    assert sum(len(r) for _, r in structs) >= 400000
    p = params(fs.LEE_RICHARDS, 30)
    split = fs.calc_coord_batch(structs, p)
    whole = eng32.calc_batch(fs.LEE_RICHARDS, structs, 1.4, 30)      # explicit context: one pass
    for a, b in zip(split, whole):
        np.testing.assert_array_equal(a.sasa, b)
    k = 77
    want = ob.oracle_calc(structs[k][0], structs[k][1], ob.LEE_RICHARDS, 1.4, 30)
    assert maxerr(split[k].sasa, want) < LR_TOL_FP32
    bad = list(structs)
    bad[100] = (np.full_like(structs[100][0], np.nan), structs[100][1])
    with pytest.raises(RuntimeError):
        fs.calc_coord_batch(bad, p)


def test_precision_from_the_environment():
    """FSB200_PRECISION=fp64 switches the drop-in entry points (no precision argument) to the all-fp64 kernels."""
    import os
    import subprocess
    import sys

    code = ("import numpy as np, freesasa_b200 as fs\n"
            "from oracle import bindings as ob\n"
            "x, r = fs.workloads.globule(4000, seed=12)\n"
            "got = fs.calc_coord(x, r, fs.Parameters(fs.LEE_RICHARDS, 1.4, 100, 40, 1)).sasa\n"
            "print(float(np.abs(got - ob.oracle_calc(x, r, ob.LEE_RICHARDS, 1.4, 40)).max()))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    errs = {}
    for mode in ("fp64", "fp32"):
        out = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, FSB200_PRECISION=mode, PYTHONPATH=root),
                             check=True, capture_output=True, text=True).stdout
        errs[mode] = float(out.strip().splitlines()[-1])
    assert errs["fp64"] < LR_TOL_FP64
    assert LR_TOL_FP64 < errs["fp32"] < LR_TOL_FP32


# ---- the fp32 tail: near-tangent circle pairs and equal arc starts -------------------------------------------------
LR_TOL_TAIL = 2e-4  # round-2 target at a million atoms, PDB-rounded coordinates, any resolution


def test_marginal_slices_are_redone_in_fp64(eng32, eng64):
    """Slices with a near-tangent pair of circles are skipped by the fp32 path and redone in fp64, and arcs with equal starts
    are ordered exactly: on coordinates rounded to the three decimals of a PDB file (where exact tangencies and ties
    happen) the fp32 engine must stay within 1.5e-4 A^2 of the reference at low resolution as well, and the certificate /
    overflow paths must be unaffected."""
    x, r = fs.workloads.globule(150000, seed=5)
    x, r = np.round(x, 3), np.round(r, 2)
    for slices in (5, 20, 100):
        want = ob.oracle_calc(x, r, ob.LEE_RICHARDS, 1.4, slices)
        got = eng32.calc(fs.LEE_RICHARDS, x, r, 1.4, slices)
        assert maxerr(got, want) < 1.5e-4, slices
        assert maxerr(eng64.calc(fs.LEE_RICHARDS, x, r, 1.4, slices), want) < LR_TOL_FP64


def test_equal_arc_starts_are_ordered_like_the_reference(eng32):
    """Atoms on an exact lattice (no jitter) give slices whose arcs share their start to the last bit; the union must not
    depend on which of them comes first (src/sasa_lr.c:367-408 sorts, ties in input order)."""
    g = np.arange(-7, 8) * 2.5
    x = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    r = np.full(len(x), 1.8)
    for slices in (4, 7, 20):
        assert maxerr(eng32.calc(0, x, r, 1.4, slices), ob.oracle_calc(x, r, 0, 1.4, slices)) < LR_TOL_TAIL, slices


# ---- full-size benchmark configurations C4, C5 and the million-atom PDB-rounded tail -------------------------------
def test_full_size_c5_one_million_atom_shell(eng32):
    """BASELINE.json configs[4]: one 1M-atom capsid-scale shell, L&R n_slices = 100, against the oracle for every atom."""
    x, r = fs.workloads.capsid(1_000_000)
    got = eng32.calc(fs.LEE_RICHARDS, x, r, 1.4, 100)
    want = ob.oracle_calc(x, r, ob.LEE_RICHARDS, 1.4, 100)
    assert maxerr(got, want) < LR_TOL_TAIL
    assert abs(got.sum() - want.sum()) < 1e-6 * want.sum()


def test_full_size_c4_batch_of_1024(eng32):
    """BASELINE.json configs[3]: 1024 independent ~5k-atom structures, L&R n_slices = 50, through the context-free batch
    entry point (overlapped sub-batches); 128 sampled structures against the oracle, every structure against a one-pass
    call on an explicit context (bit-identical)."""
    structs = fs.workloads.batch(1024, 4000, 6000, seed=0)
    got = fs.calc_batch(fs.LEE_RICHARDS, structs, 1.4, 50)
    assert [len(g) for g in got] == [len(b) for _, b in structs]
    for k in range(0, 1024, 8):
        want = ob.oracle_calc(structs[k][0], structs[k][1], ob.LEE_RICHARDS, 1.4, 50)
        assert maxerr(got[k], want) < LR_TOL_TAIL, k
    whole = eng32.calc_batch(fs.LEE_RICHARDS, structs[:200], 1.4, 50)
    for a, b in zip(got[:200], whole):
        np.testing.assert_array_equal(a, b)


def test_one_million_atoms_pdb_rounded_low_resolution(eng32):
    """The hard case for fp32: a million atoms with coordinates rounded to 3 decimals, n_slices = 20 and 5."""
    x, r = fs.workloads.globule(1_000_000, seed=5)
    x, r = np.round(x, 3), np.round(r, 2)
    for slices in (20, 5):
        got = eng32.calc(fs.LEE_RICHARDS, x, r, 1.4, slices)
        want = ob.oracle_calc(x, r, ob.LEE_RICHARDS, 1.4, slices)
        assert maxerr(got, want) < LR_TOL_TAIL, slices


def test_full_task_pool_falls_back_bit_identically():
    """The split pipeline keeps one task record per atom that has to be integrated in a pool sized for about half of the
    atoms; an atom that does not fit is integrated inside k_integrate, chunk by chunk in the same order.  With the pool
    shrunk to a few records (test hook) and with the pipeline fused, the areas must have the same bits as the normal run."""
    import os
    import subprocess
    import sys

    code = ("import numpy as np, freesasa_b200 as fs\n"
            "x, r = fs.workloads.globule(30000, seed=21)\n"
            "x, r = np.round(x, 3), np.round(r, 2)\n"
            "e = fs.Engine(0)\n"
            "a = e.calc(fs.LEE_RICHARDS, x, r, 1.4, 37)\n"
            "e.set_certificate(False)\n"
            "b = e.calc(fs.LEE_RICHARDS, x, r, 1.4, 37)\n"
            "np.save(__import__('sys').argv[1], np.stack([a, b]))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for name, env in (("normal", {}), ("tiny_pool", {"FSB200_POOL_BYTES_PER_ATOM": "16"}), ("fused", {"FSB200_PIPELINE": "fused"})):
        path = os.path.join(root, "gpurun_out", f"_pool_{name}.npy")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        subprocess.run([sys.executable, "-c", code, path], cwd=root, env=dict(os.environ, PYTHONPATH=root, **env), check=True)
        outs[name] = np.load(path)
        os.remove(path)
    np.testing.assert_array_equal(outs["normal"][0], outs["normal"][1])      # certificate on / off
    np.testing.assert_array_equal(outs["normal"], outs["tiny_pool"])         # pool full: in-kernel fallback
    np.testing.assert_array_equal(outs["normal"], outs["fused"])             # everything in one kernel
    x, r = fs.workloads.globule(30000, seed=21)
    x, r = np.round(x, 3), np.round(r, 2)
    assert maxerr(outs["normal"][0], ob.oracle_calc(x, r, ob.LEE_RICHARDS, 1.4, 37)) < LR_TOL_TAIL

"""The buried-atom certificate (freesasa_b200/csrc/integrate.cu: certify_buried).

CPU part: the geometric constant it relies on — the library's 128 probe directions (64 antipodal pairs) cover the
sphere with patches of angular radius < 12.5 degrees.  GPU part: the certificate never changes a result and only ever fires on
atoms whose reference area is exactly zero."""
import math
import os
import re

import numpy as np
import pytest

from oracle import bindings as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RHO_DEG = 12.5


def _constants():
    src = open(os.path.join(ROOT, "freesasa_b200", "csrc", "integrate.cu")).read()
    cos = float(re.search(r"kCertCos = ([0-9.]+)f", src).group(1))
    sin = float(re.search(r"kCertSin = ([0-9.]+)f", src).group(1))
    n = int(re.search(r"kCertPoints = (\d+)", open(os.path.join(ROOT, "freesasa_b200", "csrc", "engine.cuh")).read()).group(1))
    return cos, sin, n


def test_patch_radius_constant_covers_the_sphere():
    cos, sin, n = _constants()
    assert abs(cos - math.cos(math.radians(RHO_DEG))) < 1e-6 and abs(sin - math.sin(math.radians(RHO_DEG))) < 1e-6
    import ctypes

    import freesasa_b200 as fs

    pts = np.empty((n, 3))
    lib = fs._engine_lib()
    lib.fsb200_cert_directions.argtypes = [ctypes.POINTER(ctypes.c_double)]
    assert lib.fsb200_cert_directions(pts.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0  # the device's own table
    assert np.array_equal(pts[: n // 2], -pts[n // 2:])                      # antipodal pairs: one dot product per pair
    assert np.abs(np.linalg.norm(pts, axis=1) - 1).max() < 2e-7             # unit vectors up to fp32 rounding
    m = 1_000_000
    probe = ob.oracle_test_points(m)  # a dense deterministic probe set; its own covering radius is ~1.3*sqrt(4pi/m)
    best = np.full(m, -1.0)
    for i in range(0, n, 32):
        best = np.maximum(best, (probe @ pts[i : i + 32].T).max(1))
    covering = math.degrees(math.acos(best.min()))
    probe_resolution = math.degrees(1.3 * math.sqrt(4 * math.pi / m))
    assert covering + probe_resolution < RHO_DEG, (covering, probe_resolution)


@pytest.mark.gpu
def test_certificate_never_changes_a_result():
    import freesasa_b200 as fs

    on, off = fs.Engine(0), fs.Engine(0)
    off.set_certificate(False)
    f = np.load(os.path.join(ROOT, "tests", "golden", "pdb_fixtures.npz"))
    cases = [fs.workloads.globule(20000, seed=3), fs.workloads.capsid(20000, r_out=60.0, seed=2),
             (f["3bzd_trimmed_xyz"], f["3bzd_trimmed_radii"]), (f["1ubq_xyz"], f["1ubq_radii"])]
    rng = np.random.default_rng(4)
    cases.append((rng.uniform(-14, 14, (3000, 3)), rng.uniform(1.0, 2.2, 3000)))  # loose random packing: many tiny exposures
    for x, r in cases:
        for alg, res in ((fs.LEE_RICHARDS, 40), (fs.SHRAKE_RUPLEY, 300)):
            a, b = on.calc(alg, x, r, 1.4, res), off.calc(alg, x, r, 1.4, res)
            np.testing.assert_array_equal(a, b)
            assert off.stats()["n_certified"] == 0
    x, r = cases[0]
    on.calc(fs.LEE_RICHARDS, x, r, 1.4, 40)
    assert on.stats()["n_certified"] > 0.5 * len(r)  # the point of the exercise
    on.close()
    off.close()


@pytest.mark.gpu
def test_certified_atoms_have_exactly_zero_reference_area():
    import freesasa_b200 as fs

    eng = fs.Engine(0)
    rng = np.random.default_rng(11)
    structures = [fs.workloads.globule(8000, seed=9)]
    for scale in (1.04, 1.10, 1.18):  # progressively looser lattices: more and more nearly-buried atoms
        x, r = fs.workloads.globule(6000, seed=int(scale * 100))
        structures.append((x * scale, r))
    structures.append((rng.uniform(-16, 16, (5000, 3)), rng.choice([1.2, 1.6, 1.9, 2.4], 5000)))
    n_cert = 0
    for x, r in structures:
        eng.neighbour_counts(x, r, 1.4)
        cert = eng.last_certified.astype(bool)
        n_cert += int(cert.sum())
        lr = ob.oracle_calc(x, r, ob.LEE_RICHARDS, 1.4, 200)
        sr = ob.oracle_calc(x, r, ob.SHRAKE_RUPLEY, 1.4, 2000)
        assert (lr[cert] == 0.0).all() and (sr[cert] == 0.0).all()
        tiny = (lr > 0) & (lr < 1e-2)  # nearly buried atoms must never be certified
        assert not cert[tiny].any()
    assert n_cert > 5000
    eng.close()

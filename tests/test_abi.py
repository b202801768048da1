"""CPU tests: the C-ABI libraries load and export every symbol the headers declare; host-side
validation mirrors the reference (no compute without a GPU)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

import freesasa_b200 as fs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:fsb200|freesasa)_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g

    g.build()
    return fs.library_paths()


def test_engine_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built["engine"])
    names = _declared("fsb200.h")
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), name


def test_host_layer_exports_every_declared_symbol(built):
    ctypes.CDLL(built["engine"], mode=ctypes.RTLD_GLOBAL)
    lib = ctypes.CDLL(built["host"])
    for name in _declared("freesasa_b200_host.h"):
        assert hasattr(lib, name), name
    p = fs.Parameters.in_dll(lib, "freesasa_default_parameters")  # reference src/freesasa.c:38-43
    assert (p.alg, p.probe_radius, p.shrake_rupley_n_points, p.lee_richards_n_slices, p.n_threads) == (0, 1.4, 100, 20, 2)


def test_struct_layouts_match_reference_abi():
    # reference src/freesasa.h:232-238,267-272 on LP64: 32-byte parameters, 56-byte result
    assert ctypes.sizeof(fs.Parameters) == 32
    assert fs.Parameters.probe_radius.offset == 8 and fs.Parameters.n_threads.offset == 24
    from freesasa_b200 import _CResult

    assert ctypes.sizeof(_CResult) == 56 and _CResult.parameters.offset == 24


def test_shard_ranges_partition(built):
    lib = ctypes.CDLL(built["engine"])
    for n in (1, 7, 100000, 1000003):
        for k in (1, 2, 3, 8):
            edges = [(lib.fsb200_shard_begin(n, i, k), lib.fsb200_shard_end(n, i, k)) for i in range(k)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(k - 1))


def test_parameter_validation_like_reference(built, capfd):
    """reference src/sasa_lr.c:177-183, src/sasa_sr.c:188-193 and tests/test-cli.in:162-164,194-196:
    more than 16 threads and non-positive resolutions are errors (NULL result) before any compute."""
    H = fs._host_lib()
    H.freesasa_set_verbosity(0)  # FREESASA_V_NORMAL, whatever earlier tests left behind
    xyz = np.zeros(3)
    rad = np.ones(1)
    dp = ctypes.POINTER(ctypes.c_double)
    for alg in (fs.LEE_RICHARDS, fs.SHRAKE_RUPLEY):
        for bad in (fs.Parameters(alg, 1.4, 100, 20, 17), fs.Parameters(alg, 1.4, 0, 0, 1), fs.Parameters(alg, 1.4, -1, -1, 1)):
            res = H.freesasa_calc_coord(xyz.ctypes.data_as(dp), rad.ctypes.data_as(dp), 1, ctypes.byref(bad))
            assert not res
    err = capfd.readouterr().err
    assert "does not support more than 16 threads" in err
    assert "invalid resolution" in err
    H.freesasa_set_verbosity(2)  # FREESASA_V_SILENT
    bad = fs.Parameters(0, 1.4, 100, 20, 17)
    assert not H.freesasa_calc_coord(xyz.ctypes.data_as(dp), rad.ctypes.data_as(dp), 1, ctypes.byref(bad))
    assert capfd.readouterr().err == ""
    H.freesasa_set_verbosity(0)


@pytest.mark.parametrize("n", [1, 31, 100, 128, 1000, 4999])
def test_engine_test_points_are_the_reference_points_reordered(built, n):
    """The engine tests Shrake-Rupley points in patch order; the SET must be the reference's golden spiral,
    bit for bit (the count of exposed points does not depend on the order).  Host code: no GPU needed."""
    from oracle import bindings as ob

    lib = ctypes.CDLL(built["engine"])
    lib.fsb200_test_points.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    out = np.empty((n, 3))
    assert lib.fsb200_test_points(n, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
    want = ob.oracle_test_points(n)  # pinned to the reference in tests/test_oracle.py
    key = lambda a: a[np.lexsort((a[:, 0], a[:, 1], a[:, 2]))]
    np.testing.assert_array_equal(key(out), key(want))
    if n >= 64:  # consecutive groups of 32 are compact patches: far smaller than the sphere
        spans = [np.linalg.norm(out[i : i + 32] - out[i : i + 32].mean(0), axis=1).max() for i in range(0, n - 31, 32)]
        assert np.median(spans) < 3.2 * math.sqrt(4 * math.pi * 32 / n) / 2 + 0.05


def test_no_cpu_fallback(built):
    """Without a device every compute entry point must fail loudly, never fall back."""
    if fs.available():
        pytest.skip("a B200 is visible")
    with pytest.raises(RuntimeError):
        fs.Engine(0)
    H = fs._host_lib()
    H.freesasa_set_verbosity(2)
    try:
        with pytest.raises(RuntimeError, match="failed"):
            fs.calc_coord(np.zeros((2, 3)), np.ones(2))
    finally:
        H.freesasa_set_verbosity(0)


def test_product_never_touches_oracle():
    """The product tree must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "freesasa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".inc")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU", ""), os.path.join(dirpath, f)


def test_public_headers_compile_as_c99_and_cxx14(tmp_path):
    """include/freesasa.h (the forwarding header), freesasa_b200_host.h and fsb200.h are valid strict C99 and C++14."""
    import subprocess

    inc = os.path.join(ROOT, "include")
    c = os.path.join(tmp_path, "t.c")
    cxx = os.path.join(tmp_path, "t.cpp")
    body = ('#include "freesasa.h"\n#include "fsb200.h"\n'
            "int main(void) { freesasa_parameters p = freesasa_default_parameters; freesasa_nodearea a = freesasa_nodearea_null;\n"
            "  freesasa_chain_group g = {0, 0}; (void)p; (void)a; (void)g; return FREESASA_ATOM_POLAR == 1 ? 0 : 1; }\n")
    open(c, "w").write(body)
    open(cxx, "w").write(body)
    subprocess.run(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror",
                    "-I", inc, "-fsyntax-only", c], check=True)
    subprocess.run(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-I", inc,
                    "-fsyntax-only", cxx], check=True)


def test_multi_gpu_entry_points_fail_loudly_without_a_gpu():
    """fsb200_calc_multi / _lr_multi / the two-half device call: no silent fallback when no sm_100 device is visible, and the
    bookkeeping entry points (trim, statistics) work without one."""
    import numpy as np

    if fs.available():
        pytest.skip("a B200 is visible")
    x, r = fs.workloads.globule(50)
    with pytest.raises(RuntimeError, match="no sm_100 device|no CUDA|CPU path"):
        fs.calc_multi(fs.LEE_RICHARDS, [(x, r)], 1.4, 20, n_devices=2)
    with pytest.raises(RuntimeError):
        fs.calc_multi(fs.LEE_RICHARDS, [(x, r), (x, r)], 1.4, 20, n_devices=0)
    assert fs.trim() == 0
    st = fs.multi_stats()
    assert st["n_devices"] == 0 and st["total_ms"] == 0.0
    with pytest.raises(RuntimeError):
        fs.IpcBuffer(0, 1024)
    assert isinstance(np.asarray(fs.device_count()).item(), int)

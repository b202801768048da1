"""ctypes bindings for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Two libraries live under oracle/:

* ``liboracle.so``            our C restatement of the reference hot path (oracle.c)
* ``_ref/libfreesasa_ref.so`` the UNMODIFIED reference (FreeSASA 2.1.3) compiled by ``make ref``
                              from the sources under /root/reference (dev container only; the
                              built file travels to the GPU box with the snapshot)

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this
module.  Nothing under freesasa_b200/ does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libfreesasa_ref.so")

LEE_RICHARDS = 0  # enum freesasa_algorithm, reference src/freesasa.h:89-92
SHRAKE_RUPLEY = 1

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(ref: bool = True) -> None:
    """(Re)build liboracle.so and, when the reference tree is present, _ref/."""
    targets = ["oracle"] + (["ref", "patched"] if ref else [])
    subprocess.run(["make", "-C", HERE, "-s"] + targets, check=True)


def _as_f64(a, shape_last=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape_last is not None and a.ndim == 2:
        assert a.shape[1] == shape_last
    return a


def _ptr(a):
    return a.ctypes.data_as(_dp)


# ------------------------------------------------------------------------------------------------
# our restatement
# ------------------------------------------------------------------------------------------------
_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = ctypes.CDLL(ORACLE_SO)
        L.oracle_neighbours.argtypes = [_dp, _dp, ctypes.c_int, ctypes.POINTER(_ip), ctypes.POINTER(_ip)]
        L.oracle_lee_richards.argtypes = [_dp, _dp, _dp, ctypes.c_int, ctypes.c_double, ctypes.c_int]
        L.oracle_shrake_rupley.argtypes = [_dp, _dp, _dp, ctypes.c_int, ctypes.c_double, ctypes.c_int]
        L.oracle_test_points.argtypes = [ctypes.c_int, _dp]
        L.oracle_test_points.restype = None
        L.oracle_exposed_arc.argtypes = [_dp, ctypes.c_int]
        L.oracle_exposed_arc.restype = ctypes.c_double
        L.oracle_free.argtypes = [ctypes.c_void_p]
        L.oracle_free.restype = None
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        L.oracle_set_threads.restype = None
        _oracle = L
    return _oracle


def oracle_neighbours(xyz, R):
    """CSR neighbour list (start[n+1], list[start[n]]) for radii R that already include the probe."""
    L = oracle_lib()
    xyz = _as_f64(xyz).reshape(-1)
    R = _as_f64(R)
    n = R.shape[0]
    s, l = _ip(), _ip()
    if L.oracle_neighbours(_ptr(xyz), _ptr(R), n, ctypes.byref(s), ctypes.byref(l)) != 0:
        raise RuntimeError("oracle_neighbours failed")
    start = np.ctypeslib.as_array(s, shape=(n + 1,)).copy()
    lst = np.ctypeslib.as_array(l, shape=(max(int(start[n]), 1),)).copy()[: int(start[n])]
    L.oracle_free(s)
    L.oracle_free(l)
    return start, lst


def oracle_calc(xyz, radii, alg=LEE_RICHARDS, probe=1.4, resolution=20, threads=0):
    """Per-atom SASA from our restatement (double)."""
    L = oracle_lib()
    xyz = _as_f64(xyz).reshape(-1)
    radii = _as_f64(radii)
    n = radii.shape[0]
    assert xyz.shape[0] == 3 * n
    if threads:
        L.oracle_set_threads(int(threads))
    out = np.empty(n, dtype=np.float64)
    fn = L.oracle_lee_richards if alg == LEE_RICHARDS else L.oracle_shrake_rupley
    if fn(_ptr(out), _ptr(xyz), _ptr(radii), n, float(probe), int(resolution)) != 0:
        raise RuntimeError("oracle calculation failed")
    return out


def oracle_test_points(n_points):
    out = np.empty(3 * n_points, dtype=np.float64)
    oracle_lib().oracle_test_points(int(n_points), _ptr(out))
    return out.reshape(-1, 3)


def oracle_exposed_arc(arcs):
    a = np.array(arcs, dtype=np.float64).reshape(-1).copy()
    return float(oracle_lib().oracle_exposed_arc(_ptr(a), a.shape[0] // 2))


# ------------------------------------------------------------------------------------------------
# the unmodified reference
# ------------------------------------------------------------------------------------------------
class RefParameters(ctypes.Structure):
    """struct freesasa_parameters, reference src/freesasa.h:232-238."""

    _fields_ = [
        ("alg", ctypes.c_int),
        ("probe_radius", ctypes.c_double),
        ("shrake_rupley_n_points", ctypes.c_int),
        ("lee_richards_n_slices", ctypes.c_int),
        ("n_threads", ctypes.c_int),
    ]


class RefResult(ctypes.Structure):
    """struct freesasa_result, reference src/freesasa.h:267-272."""

    _fields_ = [
        ("total", ctypes.c_double),
        ("sasa", _dp),
        ("n_atoms", ctypes.c_int),
        ("parameters", RefParameters),
    ]


_ref = None


def ref_available() -> bool:
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        if not ref_available():
            raise RuntimeError(
                "oracle/_ref/libfreesasa_ref.so is missing: run `make -C oracle ref` in the "
                "dev container (needs /root/reference)"
            )
        L = ctypes.CDLL(REF_SO)
        L.freesasa_calc_coord.argtypes = [_dp, _dp, ctypes.c_int, ctypes.POINTER(RefParameters)]
        L.freesasa_calc_coord.restype = ctypes.POINTER(RefResult)
        L.freesasa_result_free.argtypes = [ctypes.POINTER(RefResult)]
        L.freesasa_result_free.restype = None
        L.freesasa_set_verbosity.argtypes = [ctypes.c_int]
        _ref = L
    return _ref


def ref_params(alg=LEE_RICHARDS, probe=1.4, resolution=20, threads=1):
    return RefParameters(int(alg), float(probe), int(resolution), int(resolution), int(threads))


def ref_calc(xyz, radii, alg=LEE_RICHARDS, probe=1.4, resolution=20, threads=1):
    """Per-atom SASA from the unmodified reference via freesasa_calc_coord (src/freesasa.c:122)."""
    L = ref_lib()
    xyz = _as_f64(xyz).reshape(-1)
    radii = _as_f64(radii)
    n = radii.shape[0]
    assert xyz.shape[0] == 3 * n and n > 0
    p = ref_params(alg, probe, resolution, threads)
    res = L.freesasa_calc_coord(_ptr(xyz), _ptr(radii), n, ctypes.byref(p))
    if not res:
        raise RuntimeError("reference freesasa_calc_coord returned NULL")
    out = np.ctypeslib.as_array(res.contents.sasa, shape=(n,)).copy()
    L.freesasa_result_free(res)
    return out


def ref_structure_from_pdb(path: str):
    """(xyz[n,3], radii[n]) exactly as the reference CLI sees a PDB file with default options
    (ProtOr radii, no hydrogens/hetatm): freesasa_structure_from_pdb(f, NULL, 0),
    reference src/structure.c:838.  Dev container only (needs the PDB file)."""
    L = ref_lib()
    libc = ctypes.CDLL(None)
    libc.fopen.restype = ctypes.c_void_p
    libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    libc.fclose.argtypes = [ctypes.c_void_p]
    L.freesasa_structure_from_pdb.restype = ctypes.c_void_p
    L.freesasa_structure_from_pdb.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    L.freesasa_structure_n.argtypes = [ctypes.c_void_p]
    L.freesasa_structure_coord_array.restype = _dp
    L.freesasa_structure_coord_array.argtypes = [ctypes.c_void_p]
    L.freesasa_structure_radius.restype = _dp
    L.freesasa_structure_radius.argtypes = [ctypes.c_void_p]
    L.freesasa_structure_free.argtypes = [ctypes.c_void_p]
    f = libc.fopen(path.encode(), b"r")
    if not f:
        raise FileNotFoundError(path)
    s = L.freesasa_structure_from_pdb(f, None, 0)
    libc.fclose(f)
    if not s:
        raise RuntimeError(f"reference could not read {path}")
    n = L.freesasa_structure_n(s)
    xyz = np.ctypeslib.as_array(L.freesasa_structure_coord_array(s), shape=(n, 3)).copy()
    rad = np.ctypeslib.as_array(L.freesasa_structure_radius(s), shape=(n,)).copy()
    L.freesasa_structure_free(s)
    return xyz, rad

"""The drop-in proof of INTEGRATION.md: the reference library with ONLY src/sasa_lr.c and src/sasa_sr.c exchanged for the
two shim files of examples/reference_patch/ (built by `make -C oracle patched` into oracle/_ref/libfreesasa_patched.so) —
its own PDB reader, classifiers, result tree and selection language, the B200 engine underneath.

CPU: the library builds, loads, reads structures with the reference's own code and refuses to compute without a GPU.
GPU (run last: the file name sorts after everything else): the patched reference against the unmodified one."""
import ctypes
import os

import numpy as np
import pytest

import freesasa_b200 as fs
from freesasa_b200 import structure as st
from freesasa_b200 import workloads as w
from oracle import bindings as ob

PATCHED = os.path.join(os.path.dirname(ob.REF_SO), "libfreesasa_patched.so")
pytestmark = pytest.mark.skipif(not (os.path.exists(PATCHED) and ob.ref_available()), reason="oracle/_ref/libfreesasa_patched.so not built")


@pytest.fixture(scope="module")
def libs():
    fs._engine_lib()  # libfsb200.so, RTLD_GLOBAL
    patched = st.StructureAPI(ctypes.CDLL(PATCHED), ob.RefResult, ob.RefParameters)
    ref = st.StructureAPI(ob.ref_lib(), ob.RefResult, ob.RefParameters)
    for api in (patched, ref):
        api.lib.freesasa_set_verbosity(2)
    return patched, ref


def test_patched_reference_loads_and_reads_with_its_own_code(libs):
    patched, ref = libs
    text = w.pdb_text(300, seed=3, chains=2, hetatm=2).encode()
    a, b = patched.from_pdb(text), ref.from_pdb(text)
    assert a.n == b.n == 300 and np.array_equal(a.xyz(), b.xyz()) and np.array_equal(a.radii(), b.radii())
    assert not hasattr(patched.lib, "freesasa_nb_new")  # src/nb.c is gone from the patched build
    bad = ob.RefParameters(fs.LEE_RICHARDS, 1.4, 100, 20, 17)  # more than 16 threads: rejected before any compute
    with pytest.raises(RuntimeError):
        a.calc(bad)
    if not fs.available():
        with pytest.raises(RuntimeError):  # no GPU: the engine says so, nothing falls back to a CPU path
            a.calc(ob.RefParameters(fs.LEE_RICHARDS, 1.4, 100, 20, 1))


@pytest.mark.gpu
@pytest.mark.parametrize("alg,res,tol", [(fs.LEE_RICHARDS, 20, 5e-4), (fs.SHRAKE_RUPLEY, 100, 1e-9)])
def test_patched_reference_matches_the_unmodified_one(libs, alg, res, tol):
    patched, ref = libs
    text = w.pdb_text(4000, seed=4, chains=3).encode()
    a, b = patched.from_pdb(text), ref.from_pdb(text)
    got, total = a.calc(ob.RefParameters(alg, 1.4, res, res, 2))
    want, want_total = b.calc(ob.RefParameters(alg, 1.4, res, res, 2))
    assert float(np.abs(got - want).max()) <= tol
    assert abs(total - want_total) <= tol * a.n
    # the reference's own tree and selection code on top of the engine's numbers
    tp, tr = st.TreeAPI(patched), st.TreeAPI(ref)
    for api in (patched, ref):
        api.lib.freesasa_calc_tree.restype = ctypes.c_void_p
        api.lib.freesasa_calc_tree.argtypes = [ctypes.c_void_p, ctypes.POINTER(ob.RefParameters), ctypes.c_char_p]
    p = ob.RefParameters(alg, 1.4, res, res, 1)
    ra, rb = patched.lib.freesasa_calc_tree(a.h, ctypes.byref(p), b"x"), ref.lib.freesasa_calc_tree(b.h, ctypes.byref(p), b"x")
    wa, wb = tp.walk(ra), tr.walk(rb)
    assert [x[:3] for x in wa] == [x[:3] for x in wb]  # same topology and names
    sa = np.uint64(next(x for x in wa if x[1] == st.NODE_STRUCTURE)[3][1]).view(np.float64)
    sb = np.uint64(next(x for x in wb if x[1] == st.NODE_STRUCTURE)[3][1]).view(np.float64)
    assert abs(float(sa) - float(sb)) <= tol * a.n
    tp.free(ra)
    tr.free(rb)

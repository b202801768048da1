"""GPU tests (run with -m gpu) for the rows around the hot path: PDB text -> structure (f-1) ->
SASA on the B200 (hot path) -> areas of residues / chains / classes (f-3), one structure at a time and as a batch (f-2).

Oracle: the compiled, unmodified reference (oracle/_ref, travels to the GPU box) run on the same text through the same
Python binding.  Bars: structures bit-identical (already asserted on CPU, re-checked here on the box's libc); per-atom
SASA within the north star's 1e-3 Å^2 for Lee-Richards (we assert 5e-4), exact for Shrake-Rupley; aggregated areas within
(number of atoms summed) x the per-atom bound, and EXACTLY the reference's aggregation of this engine's own per-atom values.
"""
import ctypes

import numpy as np
import pytest

import freesasa_b200 as fs
from freesasa_b200 import structure as st
from freesasa_b200 import workloads as w
from oracle import bindings as ob

pytestmark = pytest.mark.gpu
LR_TOL, SR_TOL = 5e-4, 1e-9


@pytest.fixture(scope="module")
def mine():
    api = st.api()
    api.lib.freesasa_set_verbosity(1)
    return api


@pytest.fixture(scope="module")
def ref():
    api = st.StructureAPI(ob.ref_lib(), ob.RefResult, ob.RefParameters)
    api.lib.freesasa_set_verbosity(1)
    return api


def areas_of(tree_api, root):
    return [(t, name, np.array([np.uint64(v).view(np.float64) for v in area[1:]])) for _, t, name, area, _, _ in tree_api.walk(root) if area]


@pytest.mark.parametrize("alg,res,tol", [(fs.LEE_RICHARDS, 20, LR_TOL), (fs.LEE_RICHARDS, 100, LR_TOL), (fs.SHRAKE_RUPLEY, 100, SR_TOL)])
@pytest.mark.parametrize("spec", [dict(n_atoms=3000, seed=1, chains=3), dict(n_atoms=20000, seed=2, chains=6, hydrogens=0.1, hetatm=20, unknown=0.02)])
def test_pdb_text_to_sasa(mine, ref, alg, res, tol, spec):
    text = w.pdb_text(**spec).encode()
    a, b = mine.from_pdb(text), ref.from_pdb(text)
    assert a.n == b.n and np.array_equal(a.xyz(), b.xyz()) and np.array_equal(a.radii(), b.radii())
    got, total = a.calc(fs.Parameters(alg, 1.4, res, res, 1))
    want, want_total = b.calc(ob.RefParameters(alg, 1.4, res, res, 1))
    assert float(np.abs(got - want).max()) <= tol
    assert total == float(np.cumsum(got)[-1])  # serial sum in atom order (src/freesasa.c:113-116); Python's sum() compensates
    assert abs(total - want_total) <= tol * a.n


def test_tree_from_pdb_text(mine, ref):
    """freesasa_calc_tree(): the areas are the reference's aggregation of the per-atom values, and close to the reference's tree."""
    text = w.pdb_text(5000, seed=3, chains=4, hetatm=5).encode()
    p = fs.Parameters(fs.LEE_RICHARDS, 1.4, 100, 20, 1)
    a, b = mine.from_pdb(text, None, st.INCLUDE_HETATM), ref.from_pdb(text, None, st.INCLUDE_HETATM)
    tm, tr = st.TreeAPI(mine), st.TreeAPI(ref)
    root = mine.lib.freesasa_calc_tree(a.h, ctypes.byref(p), b"gpu")
    assert root
    walk = tm.walk(root)
    assert [x[1] for x in walk].count(st.NODE_ATOM) == a.n
    # (1) exact: feed this engine's per-atom values to the REFERENCE's tree builder
    sasa = np.array([np.uint64(v).view(np.float64) for v in next(x for x in walk if x[1] == st.NODE_STRUCTURE)[4][6]])
    result, keep = tr.make_result(sasa, ob.RefParameters(fs.LEE_RICHARDS, 1.4, 100, 20, 1))
    ref_root = tr.init(result, b, b"gpu")
    assert [x[:4] for x in tr.walk(ref_root)] == [x[:4] for x in walk]
    tr.free(ref_root)
    # (2) within tolerance of the reference computing everything itself
    ref_root = ref.lib.freesasa_calc_tree(b.h, ctypes.byref(ob.RefParameters(fs.LEE_RICHARDS, 1.4, 100, 20, 1)), b"gpu")
    for (t1, n1, v1), (t2, n2, v2) in zip(areas_of(tm, root), areas_of(tr, ref_root)):
        assert (t1, n1) == (t2, n2)
        assert np.abs(v1 - v2).max() <= LR_TOL * max(1, a.n if t1 != st.NODE_ATOM else 1)
    tr.free(ref_root)
    tm.free(root)


@pytest.mark.parametrize("options", [st.SEPARATE_MODELS, st.SEPARATE_MODELS | st.SEPARATE_CHAINS])
def test_structure_array_as_one_batch(mine, ref, options):
    """Row f-2: every model / chain of a file in ONE device pass equals one call per structure."""
    text = w.pdb_text(1500, seed=4, chains=3, models=6).encode()
    structures = mine.array(text, None, options)
    refs = ref.array(text, None, options)
    assert len(structures) == len(refs) == (6 if options == st.SEPARATE_MODELS else 18)
    H = mine.lib
    H.freesasa_calc_structure_batch.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(fs.Parameters),
                                                ctypes.POINTER(ctypes.POINTER(fs._CResult))]
    n = len(structures)
    handles = (ctypes.c_void_p * n)(*[s.h for s in structures])
    results = (ctypes.POINTER(fs._CResult) * n)()
    p = fs.Parameters(fs.SHRAKE_RUPLEY, 1.4, 200, 20, 1)
    assert H.freesasa_calc_structure_batch(n, handles, ctypes.byref(p), results) == 0
    for k, (s, r) in enumerate(zip(structures, refs)):
        got = np.ctypeslib.as_array(results[k].contents.sasa, shape=(s.n,)).copy()
        want, _ = r.calc(ob.RefParameters(fs.SHRAKE_RUPLEY, 1.4, 200, 20, 1))
        single, _ = s.calc(p)
        assert np.array_equal(got, single)
        assert float(np.abs(got - want).max()) <= SR_TOL
        H.freesasa_result_free(results[k])


def test_calc_tree_batch_equals_one_tree_per_structure(mine):
    """Additive freesasa_calc_tree_batch(): one device pass + trees built concurrently == freesasa_calc_tree() per structure."""
    text = w.pdb_text(1200, seed=6, chains=3, models=9, hetatm=2).encode()
    structures = mine.array(text, None, st.SEPARATE_MODELS | st.INCLUDE_HETATM)
    n = len(structures)
    assert n == 9
    H, tree = mine.lib, st.TreeAPI(mine)
    H.freesasa_calc_tree_batch.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(fs.Parameters),
                                           ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_void_p)]
    p = fs.Parameters(fs.SHRAKE_RUPLEY, 1.4, 150, 20, 1)
    handles = (ctypes.c_void_p * n)(*[s.h for s in structures])
    names = (ctypes.c_char_p * n)(*[b"model-%d" % k for k in range(n)])
    trees = (ctypes.c_void_p * n)()
    assert H.freesasa_calc_tree_batch(n, handles, ctypes.byref(p), names, trees) == 0
    for k in range(n):
        single = H.freesasa_calc_tree(structures[k].h, ctypes.byref(p), names[k])
        assert tree.walk(trees[k]) == tree.walk(single)
        tree.free(single)
        tree.free(trees[k])
    assert H.freesasa_calc_tree_batch(n, handles, ctypes.byref(p), None, trees) == 0  # names are optional
    assert tree.walk(trees[0])[1][2] is None
    for k in range(n):
        tree.free(trees[k])

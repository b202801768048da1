"""CPU tests for scope row f-3: per-atom SASA -> areas of residues, chains, structures and classes.

The sums are floating point, but the order of every addition is the reference's (src/node.c:148-176,718-777), so the
bar is bit-exact: the whole tree — topology, names, properties, every component of every area — is compared with the
tree the compiled reference builds from the same structure and the same per-atom values.  No GPU is involved: the
per-atom values are seeded random numbers or the CPU restatement's SASA.
"""
import ctypes

import numpy as np
import pytest

from freesasa_b200 import structure as st
from freesasa_b200 import workloads as w
from oracle import bindings as ob

needs_ref = pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def both():
    mine = st.api()
    ref = st.StructureAPI(ob.ref_lib(), ob.RefResult, ob.RefParameters)
    for api in (mine, ref):
        api.lib.freesasa_set_verbosity(2)
    yield (mine, st.TreeAPI(mine)), (ref, st.TreeAPI(ref))
    mine.lib.freesasa_set_verbosity(0)


CASES = [
    dict(n_atoms=300, seed=1),
    dict(n_atoms=800, seed=2, chains=4, hydrogens=0.2, hetatm=3, unknown=0.1),
    dict(n_atoms=500, seed=3, chains=2, altloc=0.1),
    dict(n_atoms=2500, seed=4, chains=7),
]


def trees_for(both, text, options, sasa_seed, name=b"test"):
    out = []
    for api, tree in both:
        s = api.from_pdb(text, None, options)
        rng = np.random.default_rng(sasa_seed)
        sasa = rng.uniform(0, 60, size=s.n) * (rng.random(s.n) < 0.6)  # many exact zeros, like buried atoms
        result, keep = tree.make_result(sasa)
        root = tree.init(result, s, name)
        assert root
        out.append((tree.walk(root), tree.classes(s, result)))
        assert tree.free(root) == 0
        del keep
    return out


@needs_ref
@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("options", [0, st.INCLUDE_HETATM | st.INCLUDE_HYDROGEN])
def test_tree_is_bit_identical(both, case, options):
    text = w.pdb_text(**CASES[case]).encode()
    (mine_walk, mine_classes), (ref_walk, ref_classes) = trees_for(both, text, options, 100 + case)
    assert len(mine_walk) == len(ref_walk)
    for a, b in zip(mine_walk, ref_walk):
        assert a == b
    assert mine_classes == ref_classes


@needs_ref
def test_tree_with_oracle_sasa_and_odd_chains(both):
    """Chains that come back (A, B, A), blank chains, insertion codes; per-atom values from the CPU restatement."""
    from tests.test_ingest import EDGE_TEXTS

    for tag in ("chain_returns", "blank_chain", "insertion_codes", "same_number_new_chain", "altloc_runs", "nucleic"):
        text = EDGE_TEXTS[tag].encode()
        walks = []
        for api, tree in both:
            s = api.from_pdb(text)
            sasa = ob.oracle_calc(s.xyz(), s.radii(), ob.LEE_RICHARDS, 1.4, 20)
            result, keep = tree.make_result(sasa)
            root = tree.init(result, s, tag.encode())
            walks.append(tree.walk(root))
            tree.free(root)
        assert walks[0] == walks[1], tag


@needs_ref
def test_add_result_and_join(both):
    text1 = w.pdb_text(200, seed=11, chains=2).encode()
    text2 = w.pdb_text(150, seed=12).encode()
    walks = []
    for api, tree in both:
        L = api.lib
        s1, s2 = api.from_pdb(text1), api.from_pdb(text2)
        r1, k1 = tree.make_result(np.linspace(0, 30, s1.n))
        r2, k2 = tree.make_result(np.linspace(5, 50, s2.n))
        root = L.freesasa_tree_new()
        assert L.freesasa_tree_add_result(root, ctypes.byref(r1), s1.h, b"first") == 0
        assert L.freesasa_tree_add_result(root, ctypes.byref(r2), s2.h, b"second") == 0  # prepended (src/node.c:467)
        other = ctypes.c_void_p(tree.init(r2, s2, None))
        assert L.freesasa_tree_join(root, ctypes.byref(other)) == 0 and not other.value
        s1.free()
        s2.free()  # the tree owns copies of everything it needs
        walks.append(tree.walk(root))
        assert L.freesasa_node_free(L.freesasa_node_children(L.freesasa_node_children(root))) == -1  # not a root
        tree.free(root)
    assert walks[0] == walks[1]
    assert [x[2] for x in walks[0] if x[1] == st.NODE_RESULT] == [b"second", b"first", None]


def test_known_sums():
    """Hand-checkable: two residues, areas 1..5, class and backbone split (src/node.c:718-746)."""
    mine = st.api()
    tree = st.TreeAPI(mine)
    s = mine.new()
    for k, (name, res, num) in enumerate([(b" N  ", b"ALA", b"   1 "), (b" CA ", b"ALA", b"   1 "), (b" CB ", b"ALA", b"   1 "),
                                          (b" O  ", b"GLY", b"   2 "), (b" XX ", b"GLY", b"   2 ")]):
        s.add_atom(name, res, num, b"A", 4.0 * k, 0.0, 0.0)
    result, keep = tree.make_result(np.array([1.0, 2.0, 3.0, 4.0, 5.0]))
    root = tree.init(result, s, b"x")
    walk = tree.walk(root)
    f = lambda bits: np.uint64(bits).view(np.float64).item()  # noqa: E731
    areas = {(t, name): tuple(f(v) for v in area[1:]) for _, t, name, area, _, _ in walk if area}
    #                                     total main side polar apolar unknown
    assert areas[(st.NODE_RESIDUE, b"ALA")] == (6.0, 3.0, 3.0, 1.0, 5.0, 0.0)
    assert areas[(st.NODE_RESIDUE, b"GLY")] == (9.0, 4.0, 5.0, 4.0, 0.0, 5.0)
    assert areas[(st.NODE_CHAIN, b"A")] == (15.0, 7.0, 8.0, 5.0, 5.0, 5.0)
    assert areas[(st.NODE_STRUCTURE, b"A")] == (15.0, 7.0, 8.0, 5.0, 5.0, 5.0)
    assert tree.classes(s, result)[0] == b"whole-structure"
    tree.free(root)


REAL = ["1ubq", "1d3z", "2jo4", "3bzd_trimmed", "2isk", "1sui"]


@needs_ref
@pytest.mark.parametrize("name", REAL)
def test_reference_test_files(both, name):
    """The reference's own test structures (dev container only): whole tree, with and without hetero atoms/hydrogens."""
    import os

    path = f"/root/reference/tests/data/{name}.pdb"
    if not os.path.exists(path):
        pytest.skip("reference test data not present")
    with open(path, "rb") as f:
        text = f.read()
    for options in (0, st.INCLUDE_HETATM | st.INCLUDE_HYDROGEN | st.JOIN_MODELS):
        (mine_walk, mine_classes), (ref_walk, ref_classes) = trees_for(both, text, options, 7, name.encode())
        assert mine_walk == ref_walk
        assert mine_classes == ref_classes


@needs_ref
def test_write_pdb_is_byte_identical(both, tmp_path):
    """Row f-4 writer (src/pdb.c:284-375): same bytes as the reference apart from the program name in the first REMARK."""
    import os

    text = w.pdb_text(400, seed=21, chains=3, models=1, hetatm=3, altloc=0.1).encode()
    outputs = []
    for api, tree in both:
        L = api.lib
        L.freesasa_write_pdb.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        s = api.from_pdb(text, None, st.INCLUDE_HETATM)
        rng = np.random.default_rng(5)
        result, keep = tree.make_result(rng.uniform(0, 1200, size=s.n) * (rng.random(s.n) < 0.5))
        root = tree.init(result, s, b"w")
        path = os.path.join(tmp_path, "out.pdb")
        fp = st._libc.fopen(path.encode(), b"w")
        assert L.freesasa_write_pdb(fp, root) == 0
        st._libc.fclose(fp)
        tree.free(root)
        outputs.append(open(path, "rb").read().split(b"\n"))
    assert outputs[0][0].startswith(b"REMARK 999 This PDB file was generated by")
    assert outputs[0][1:] == outputs[1][1:]
    assert len(outputs[0]) > 400


@needs_ref
def test_threaded_tree_build_is_identical(both):
    """Structures of >= 20 000 atoms are built by several threads (areas.c: build_part); forced on for small ones through
    the test hook.  Every thread count gives the reference's tree, bit for bit."""
    import os

    from tests.test_ingest import EDGE_TEXTS

    texts = [w.pdb_text(1200, seed=31, chains=5, hetatm=4, unknown=0.1).encode(), EDGE_TEXTS["chain_returns"].encode() * 7,
             (EDGE_TEXTS["nucleic"] * 9).encode()]
    old = {k: os.environ.get(k) for k in ("FREESASA_B200_PARALLEL_MIN_ATOMS", "FREESASA_B200_THREADS")}
    try:
        os.environ["FREESASA_B200_PARALLEL_MIN_ATOMS"] = "1"
        for threads in (2, 3, 8, 16):
            os.environ["FREESASA_B200_THREADS"] = str(threads)
            for k, text in enumerate(texts):
                (mine_walk, mine_classes), (ref_walk, ref_classes) = trees_for(both, text, st.INCLUDE_HETATM, 300 + k)
                assert mine_walk == ref_walk
                assert mine_classes == ref_classes
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@needs_ref
def test_trees_of_a_batch_built_concurrently(both):
    """The tree half of the additive freesasa_calc_tree_batch(): many (structure, result) pairs -> trees on several threads;
    each equals the reference's freesasa_tree_init() of the same pair.  Without a GPU the full call must fail loudly."""
    import os

    import freesasa_b200 as fs

    (mine, tm), (ref, tr) = both
    text = w.pdb_text(700, seed=41, chains=2, models=7, hetatm=2).encode()
    sm, sr = mine.array(text, None, st.SEPARATE_MODELS), ref.array(text, None, st.SEPARATE_MODELS)
    n = len(sm)
    rng = np.random.default_rng(8)
    sasa = [rng.uniform(0, 50, size=s.n) for s in sm]
    res_m = [tm.make_result(a) for a in sasa]
    res_r = [tr.make_result(a) for a in sasa]
    H = mine.lib
    H.fsb_trees_from_results.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.POINTER(mine.Result)),
                                         ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_void_p)]
    handles = (ctypes.c_void_p * n)(*[s.h for s in sm])
    results = (ctypes.POINTER(mine.Result) * n)(*[ctypes.pointer(r[0]) for r in res_m])
    names = (ctypes.c_char_p * n)(*[b"m%d" % k for k in range(n)])
    trees = (ctypes.c_void_p * n)()
    old = os.environ.get("FREESASA_B200_THREADS")
    try:
        for threads in ("1", "3", "16"):
            os.environ["FREESASA_B200_THREADS"] = threads
            assert H.fsb_trees_from_results(n, handles, results, names, trees) == 0
            for k in range(n):
                want = tr.init(res_r[k][0], sr[k], b"m%d" % k)
                assert tm.walk(trees[k]) == tr.walk(want)
                tr.free(want)
                tm.free(trees[k])
    finally:
        if old is None:
            os.environ.pop("FREESASA_B200_THREADS", None)
        else:
            os.environ["FREESASA_B200_THREADS"] = old
    if not fs.available():
        H.freesasa_calc_tree_batch.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(fs.Parameters),
                                               ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_void_p)]
        p = fs.default_parameters()
        assert H.freesasa_calc_tree_batch(n, handles, ctypes.byref(p), names, trees) == -1
        assert not any(trees[k] for k in range(n))


@needs_ref
def test_write_pdb_without_pdb_lines_fails_in_both(both, tmp_path):
    """Atoms added by hand carry no PDB record: freesasa_write_pdb() reports failure (src/pdb.c:318-320); trees with several
    results (freesasa_tree_add_result twice) are written result by result."""
    import os

    outputs = []
    for api, tree in both:
        L = api.lib
        L.freesasa_write_pdb.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        s = api.new()
        for k, name in enumerate((b" N  ", b" CA ", b" C  ")):
            s.add_atom(name, b"ALA", b"   1 ", b"A", 1.5 * k, 0.0, 0.0)
        result, keep = tree.make_result(np.array([1.0, 2.0, 3.0]))
        root = tree.init(result, s, b"hand")
        fp = st._libc.fopen(os.path.join(tmp_path, "x.pdb").encode(), b"w")
        rc_fail = L.freesasa_write_pdb(fp, root)
        st._libc.fclose(fp)
        tree.free(root)
        text = w.pdb_text(60, seed=2, chains=2).encode()
        s1, s2 = api.from_pdb(text), api.from_pdb(w.pdb_text(40, seed=3).encode())
        r1, k1 = tree.make_result(np.linspace(0, 9, s1.n))
        r2, k2 = tree.make_result(np.linspace(3, 5, s2.n))
        root = tree.init(r1, s1, b"one")
        assert L.freesasa_tree_add_result(root, ctypes.byref(r2), s2.h, b"two") == 0
        path = os.path.join(tmp_path, "y.pdb")
        fp = st._libc.fopen(path.encode(), b"w")
        rc_ok = L.freesasa_write_pdb(fp, root)
        st._libc.fclose(fp)
        tree.free(root)
        outputs.append((rc_fail, rc_ok, open(path, "rb").read().split(b"\n")[1:]))
    assert outputs[0] == outputs[1]
    assert outputs[0][0] == -1 and outputs[0][1] == 0
    assert sum(ln.startswith(b"MODEL") for ln in outputs[0][2]) == 2


def _written(api, tree, root, which, tmp_path):
    import os

    f = getattr(api.lib, "freesasa_write_" + which)
    f.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    path = os.path.join(tmp_path, which + ".txt")
    fp = st._libc.fopen(path.encode(), b"w")
    assert f(fp, root) == 0
    st._libc.fclose(fp)
    return open(path, "rb").read()


def test_published_per_residue_goldens(both, pdb_fixtures, tmp_path):
    """SURVEY.md §8(c) pins: the reference's own golden files for `freesasa -S --format=res|seq < 1ubq.pdb`
    (tests/data/restype.reference, seq.reference; tests/test-cli.in:286-298).  Per-atom Shrake-Rupley values from the
    committed fixture (generated by the unmodified reference), structure labels from the reference's 1ubq.pdb (dev container
    only), residue sums + writers from this repo: the bytes must be the published ones."""
    import os

    data = "/root/reference/tests/data"
    if not os.path.exists(os.path.join(data, "1ubq.pdb")):
        pytest.skip("reference test data not present")
    (mine, tm), _ = both
    s = mine.from_pdb_path(os.path.join(data, "1ubq.pdb"))
    np.testing.assert_array_equal(s.xyz(), pdb_fixtures["1ubq_xyz"])
    np.testing.assert_array_equal(s.radii(), pdb_fixtures["1ubq_radii"])
    result, keep = tm.make_result(pdb_fixtures["1ubq_sr100"])
    root = tm.init(result, s, b"stdin")
    assert _written(mine, tm, root, "res", tmp_path) == open(os.path.join(data, "restype.reference"), "rb").read()
    assert _written(mine, tm, root, "seq", tmp_path) == open(os.path.join(data, "seq.reference"), "rb").read()
    # `freesasa -S --format=pdb < 1ubq.pdb | grep -v REMARK` == tests/data/1ubq.B.pdb (tests/test-cli.in:299-303)
    pdb = b"".join(ln + b"\n" for ln in _written(mine, tm, root, "pdb", tmp_path).split(b"\n")[:-1] if not ln.startswith(b"REMARK"))
    assert pdb == open(os.path.join(data, "1ubq.B.pdb"), "rb").read()
    tm.free(root)


@needs_ref
def test_res_and_seq_writers_are_byte_identical(both, tmp_path):
    text = w.pdb_text(900, seed=23, chains=3, hetatm=4, unknown=0.1).encode() + (
        w.pdb_atom_line(9001, "P", "  A", "D", 1, 0.0, 0.0, 9.0, "P") + "\n" + w.pdb_atom_line(9002, "C1'", " DG", "D", 2, 2.0, 0.0, 9.0, "C") + "\n").encode()
    outs = []
    for api, tree in both:
        s = api.from_pdb(text, None, st.INCLUDE_HETATM)
        rng = np.random.default_rng(6)
        result, keep = tree.make_result(rng.uniform(0, 300, size=s.n) * (rng.random(s.n) < 0.6))
        root = tree.init(result, s, b"a name")
        outs.append((_written(api, tree, root, "res", tmp_path), _written(api, tree, root, "seq", tmp_path)))
        tree.free(root)
    assert outs[0] == outs[1]
    assert b"RES HOH" not in outs[0][0] and b"RES UNK" in outs[0][0] and b"RES DG " in outs[0][0]
    for name in (b"ALA", b"  A", b"DA ", b" dg", b"XYZ", b"", b"hoh", b"N", b"GLX"):
        assert both[0][0].lib.freesasa_classify_residue(name) == both[1][0].lib.freesasa_classify_residue(name)


@pytest.mark.parametrize("name,key,total,polar,apolar", [
    ("1ubq", "lr20", 4804.055641, 2504.217302, 2299.838339),      # reference tests/test_freesasa.c:155-166
    ("1ubq", "sr100", 4834.716265, 2515.821238, 2318.895027),     # :168-178
    ("3bzd_trimmed", "sr100", 16133.867124, 7432.608118, 8701.259006),   # :305-307
])
def test_published_class_sums(both, pdb_fixtures, name, key, total, polar, apolar):
    """freesasa_result_classes() (src/classifier.c:829-838) against the polar / apolar totals the reference's own unit tests
    assert (tolerance 1e-5 there): per-atom values from the committed reference fixtures, classes from this repo's reader."""
    import os

    path = f"/root/reference/tests/data/{name}.pdb"
    if not os.path.exists(path):
        pytest.skip("reference test data not present")
    (mine, tm), _ = both
    s = mine.from_pdb_path(path)
    result, keep = tm.make_result(pdb_fixtures[f"{name}_{key}"])
    got = tm.classes(s, result)
    f = lambda bits: np.uint64(bits).view(np.float64).item()  # noqa: E731
    assert got[0] == b"whole-structure"
    assert abs(f(got[1]) - total) < 1e-5 and abs(f(got[4]) - polar) < 1e-5 and abs(f(got[5]) - apolar) < 1e-5
    assert f(got[6]) == 0.0  # nothing unknown


@pytest.mark.parametrize("classifier", ["protor", "naccess"])
def test_published_relative_areas_of_tripeptides(both, classifier):
    """SURVEY.md §8(c) pin (tests/test-cli.in:309-322): for each GLY-X-GLY tripeptide of tests/data/rsa/, Lee-Richards with
    1000 slices and the ProtOr or NACCESS radii gives residue X exactly 100.0 % of the classifier's reference areas (all
    atoms, side chain, main chain, apolar, polar — printed with one decimal).  Checks the residue reference tables, the
    classes and the tree's residue sums of this repo; the per-atom areas come from the CPU restatement."""
    import glob
    import os

    files = sorted(glob.glob("/root/reference/tests/data/rsa/*.pdb"))
    if not files:
        pytest.skip("reference test data not present")
    (mine, tm), _ = both
    L = mine.lib
    for path in files:
        s = mine.from_pdb_path(path, mine.classifier(classifier))
        sasa = ob.oracle_calc(s.xyz(), s.radii(), ob.LEE_RICHARDS, 1.4, 1000)
        result, keep = tm.make_result(sasa)
        root = tm.init(result, s, b"rsa")
        chain = L.freesasa_node_children(L.freesasa_node_children(L.freesasa_node_children(root)))
        residue = L.freesasa_node_next(L.freesasa_node_children(chain))           # residue 2 of the tripeptide
        assert L.freesasa_node_name(residue).decode() == os.path.basename(path)[:3]
        area, ref = L.freesasa_node_area(residue).contents, L.freesasa_node_residue_reference(residue).contents
        for field in ("total", "side_chain", "main_chain", "apolar", "polar"):
            absolute, reference = getattr(area, field), getattr(ref, field)
            if reference == 0.0:                                                    # GLY has no side chain: "N/A"
                assert field == "side_chain" and "GLY" in path
                continue
            assert "%.1f" % (100.0 * absolute / reference) == "100.0", (path, field, absolute, reference)
        tm.free(root)

"""CPU tests for scope row f-1: PDB text -> structure and the classifiers.

Everything here is byte / index / IEEE-bit work, so the bar is EXACT equality with the reference: the same atoms
in the same order, bit-identical coordinates and radii, the same labels, residues, chains, model numbers and the same
NULL-or-not outcome, for every option combination.  Three anchors:

  * the compiled, unmodified reference (oracle/_ref) driven through the very same Python binding
    (freesasa_b200.structure.StructureAPI) on synthetic PDB text and, in the dev container, on the reference's own
    test files (reference tests/test_structure.c reads the same files);
  * known answers the reference's tests assert (tests/test_structure.c, tests/test_classifier.c, tests/test_pdb.c);
  * committed digests of what the reference produced for seeded synthetic files (tests/golden/ingest_golden.json,
    written by tests/golden/make_ingest_golden.py), which hold even where oracle/_ref is absent.
"""
import ctypes
import glob
import hashlib
import json
import os

import numpy as np
import pytest

from freesasa_b200 import structure as st
from freesasa_b200 import workloads as w
from oracle import bindings as ob

REF_DATA = "/root/reference/tests/data"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ingest_golden.json")
OPTION_SETS = [0, st.INCLUDE_HETATM, st.INCLUDE_HYDROGEN, st.INCLUDE_HETATM | st.INCLUDE_HYDROGEN, st.JOIN_MODELS,
               st.INCLUDE_HETATM | st.INCLUDE_HYDROGEN | st.JOIN_MODELS, st.SKIP_UNKNOWN, st.HALT_AT_UNKNOWN,
               st.SKIP_UNKNOWN | st.HALT_AT_UNKNOWN, st.RADIUS_FROM_OCCUPANCY, st.INCLUDE_HETATM | st.SKIP_UNKNOWN]

needs_ref = pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def mine():
    api = st.api()
    api.lib.freesasa_set_verbosity(2)  # silent: the tests provoke errors on purpose
    yield api
    api.lib.freesasa_set_verbosity(0)


@pytest.fixture(scope="module")
def ref():
    api = st.StructureAPI(ob.ref_lib(), ob.RefResult, ob.RefParameters)
    api.lib.freesasa_set_verbosity(2)
    return api


def snapshot(s):
    """Everything observable about a structure, in comparable form (floats as bit patterns)."""
    if s is None:
        return None
    return {
        "n": s.n, "n_residues": s.n_residues, "n_chains": s.n_chains, "model": s.model,
        "chain_labels": s.chain_labels, "classifier": s.classifier_name,
        "xyz": s.xyz().view(np.uint64).tolist(), "radii": s.radii().view(np.uint64).tolist(),
        "atoms": [(a[:6] + (np.float64(a[6]).view(np.uint64).item(), a[7])) for a in s.atoms()],
        "residues": s.residues(), "chains": s.chains(),
    }


def digest(snap) -> str:
    return hashlib.sha256(repr(snap).encode()).hexdigest()


def same_structure(mine, ref, text, classifier=None, options=0, ignore=()):
    a = mine.from_pdb(text, mine.classifier(classifier) if classifier else None, options)
    b = ref.from_pdb(text, ref.classifier(classifier) if classifier else None, options)
    sa, sb = snapshot(a), snapshot(b)
    assert (sa is None) == (sb is None)
    if sa is not None:
        for key in sb:
            if key not in ignore:
                assert sa[key] == sb[key], key
    return sa


# ---- synthetic files, every option set ------------------------------------------------------------------------
SYNTHETIC = [
    dict(n_atoms=300, seed=1),
    dict(n_atoms=411, seed=2, chains=3, hydrogens=0.3, hetatm=4),
    dict(n_atoms=350, seed=3, chains=2, altloc=0.15, unknown=0.1),
    dict(n_atoms=260, seed=4, models=3, hydrogens=0.1, hetatm=2, altloc=0.05),
    dict(n_atoms=280, seed=5, element_column=False, hydrogens=0.2),
    dict(n_atoms=290, seed=6, newline="\r\n", chains=2, unknown=0.2, hetatm=3),
    dict(n_atoms=2200, seed=7, chains=5, offset=(900.0, -950.0, 1234.5)),
]


@needs_ref
@pytest.mark.parametrize("case", range(len(SYNTHETIC)))
@pytest.mark.parametrize("options", OPTION_SETS)
def test_synthetic_matches_reference(mine, ref, case, options):
    text = w.pdb_text(**SYNTHETIC[case]).encode()
    same_structure(mine, ref, text, None, options)


@needs_ref
@pytest.mark.parametrize("classifier", ["protor", "naccess", "oons"])
def test_builtin_classifiers_on_a_structure(mine, ref, classifier):
    text = w.pdb_text(500, seed=11, chains=2, hydrogens=0.2, hetatm=3, unknown=0.1).encode()
    for options in (0, st.INCLUDE_HETATM | st.INCLUDE_HYDROGEN):
        snap = same_structure(mine, ref, text, classifier, options)
        assert snap["n"] > 0


def test_committed_digests(mine):
    """The reference's output for seeded synthetic files, as digests committed by make_ingest_golden.py."""
    with open(GOLDEN) as f:
        gold = json.load(f)
    assert len(gold["cases"]) >= 20
    for case in gold["cases"]:
        text = w.pdb_text(**case["pdb_text"]).encode()
        assert hashlib.sha256(text).hexdigest() == case["text_sha256"], "generator drifted: regenerate the golden file"
        s = mine.from_pdb(text, mine.classifier(case["classifier"]) if case["classifier"] else None, case["options"])
        assert (s is None) == (case["digest"] is None)
        if s is not None:
            assert s.n == case["n_atoms"]
            assert digest(snapshot(s)) == case["digest"]


# ---- the reference's own test files (dev container only) ---------------------------------------------------------
REAL = sorted(glob.glob(os.path.join(REF_DATA, "*.pdb")) + glob.glob(os.path.join(REF_DATA, "rsa", "*.pdb")))


@needs_ref
@pytest.mark.skipif(not REAL, reason="reference test data not present")
@pytest.mark.parametrize("path", REAL, ids=[os.path.basename(p) for p in REAL])
def test_reference_test_files(mine, ref, path):
    with open(path, "rb") as f:
        text = f.read()
    for options in OPTION_SETS:
        same_structure(mine, ref, text, None, options)
    same_structure(mine, ref, text, "naccess", st.INCLUDE_HETATM | st.INCLUDE_HYDROGEN)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DATA, "1ubq.pdb")), reason="reference test data not present")
def test_known_answers_1ubq(mine):
    """reference tests/test_structure.c:160-200 (1ubq: 602 atoms, 76 residues, chain 'A', first/last atoms)."""
    s = mine.from_pdb_path(os.path.join(REF_DATA, "1ubq.pdb"))
    assert (s.n, s.n_residues, s.n_chains, s.chain_labels) == (602, 76, 1, b"A")
    atoms = s.atoms()
    assert atoms[0][:4] == (b" N  ", b"MET", b"   1 ", b"A")
    assert atoms[601][:4] == (b" OXT", b"GLY", b"  76 ", b"A")
    s_h = mine.from_pdb_path(os.path.join(REF_DATA, "1ubq.pdb"), None, st.INCLUDE_HETATM)
    assert s_h.n == 660  # 58 waters, tests/test_structure.c:203-206
    np.testing.assert_array_equal(s.xyz()[0], [27.340, 24.430, 2.614])


# ---- edge cases --------------------------------------------------------------------------------------------
def atom(serial=1, name="CA", res="ALA", chain="A", seq=1, xyz=(1.0, 2.0, 3.0), element="C", **kw):
    return w.pdb_atom_line(serial, name, res, chain, seq, *xyz, element, **kw)


EDGE_TEXTS = {
    "empty": "",
    "no_atoms": "HEADER nothing here\nREMARK\nEND\n",
    "only_hydrogens": atom(name="H", element="H") + "\n",
    "no_trailing_newline": atom() + "\n" + atom(2, "CB", xyz=(2.5, 2.0, 3.0)),
    "short_line_fails": atom() + "\n" + atom(2, "CB")[:50] + "\n",
    "line_of_12_chars": atom() + "\nATOM      2 \n" + atom(3, "CB") + "\n",
    "line_of_16_chars": atom() + "\nATOM      2  CB \n",
    "truncated_at_54": atom()[:54] + "\n" + atom(2, "CB")[:54] + "\n",
    "truncated_at_78": atom()[:78] + "\n" + atom(2, "HB", element="H")[:78] + "\n",
    "truncated_at_77": atom()[:77] + "\n" + atom(2, "HB", element="H")[:77] + "\n",
    "long_lines": atom() + " " * 30 + "\n" + atom(2, "CB") + "x" * 60 + "\n" + atom(3, "C") + "\n",
    "very_long_line": "REMARK " + "y" * 300 + "\n" + atom() + "\n" + "REMARK " + "z" * 111 + "ATOM  bogus\n" + atom(2, "CB") + "\n",
    "atom_prefix_only": "ATOMS ARE FUN, THIS LINE IS LONG ENOUGH TO PASS THE LENGTH CHECK 1.0 2.0 3.0       \n" + atom() + "\n",
    "altloc_runs": "\n".join([atom(1, "N", alt="A"), atom(2, "N", alt="B"), atom(3, "CA", alt="B"), atom(4, "C"), atom(5, "O", alt="B"),
                              atom(6, "CB", alt="A"), atom(7, "CB", alt="B")]) + "\n",
    "altloc_hydrogen_between": "\n".join([atom(1, "N", alt="A"), atom(2, "H", element="H"), atom(3, "CA", alt="B"), atom(4, "C", alt="A")]) + "\n",
    "chain_returns": "\n".join([atom(1, "N", chain="A"), atom(2, "CA", chain="B", seq=2), atom(3, "C", chain="A", seq=3), atom(4, "O", chain="A", seq=3)]) + "\n",
    "blank_chain": "\n".join([atom(1, "N", chain=" "), atom(2, "CA", chain=" ")]) + "\n",
    "insertion_codes": "\n".join([atom(1, "N", seq=10), atom(2, "N", seq=10, icode="A"), atom(3, "N", seq=10, icode="B"), atom(4, "CA", seq=10, icode="B")]) + "\n",
    "same_number_new_chain": "\n".join([atom(1, "N", chain="A", seq=5), atom(2, "N", chain="B", seq=5)]) + "\n",
    "model_numbers": "MODEL       17\n" + atom() + "\nENDMDL\nMODEL       18\n" + atom(2, "CB") + "\nENDMDL\n",
    "model_short": "MODEL\n" + atom() + "\nENDMDL\n" + atom(2, "CB") + "\n",
    "endmdl_first": "ENDMDL\n" + atom() + "\n",
    "unknown_element": atom(1, "XX", res="ZZZ", element="Xx") + "\n" + atom(2, "FE", res="HEM", element="FE") + "\n",
    "four_letter_name": atom(1, "HG11", res="VAL", element="")[:76] + "\n" + atom(2, "CG1", res="VAL") + "\n",
    "selenomet": atom(1, "SE", res="MSE", element="SE") + "\n" + atom(2, "CA", res="MSE") + "\n",
    "nucleic": atom(1, "P", res="  A", element="P") + "\n" + atom(2, "OP1", res=" DA", element="O") + "\n" + atom(3, "C5'", res="  U") + "\n",
    "tabs_in_labels": atom(1, "CA").replace(" CA ", "\tCA ") + "\n",
    "embedded_nul": atom() + "\n" + atom(2, "CB")[:40] + "\0" + atom(2, "CB")[41:] + "\n" + atom(3, "C") + "\n",
    "nul_after_coordinates": atom() + "\n" + atom(2, "CB")[:60] + "\0" + atom(2, "CB")[61:] + "\n" + atom(3, "C") + "\n",
}
# coordinate sections that leave the %8.3f layout: the reference reads them as whitespace-separated tokens (sscanf)
for tag, section in {
    "coords_free_format": " 1.5 -2.25 3       ", "coords_exponent": " 1.5e1  -2.5E-1 3e0   ", "coords_plus": "  +1.500  +2.000  -3.000",
    "coords_merged_minus": "1234.567-123.456-999.999", "coords_merged_plus": "1234.5671234.5671234.567", "coords_two_numbers": "   1.000   2.000        ",
    "coords_garbage": "   1.000   abc     3.000", "coords_nan": "     nan     inf   1.000", "coords_hex": "   0x1p3   1.000   2.000",
    "coords_long_mantissa": "1.2345678901234567 2 3  ", "coords_dot_only": "       .   1.000   2.000", "coords_leading_dot": "    .500   -.250  -0.000",
    "coords_trailing_dot": "      5.     -6.      7.", "coords_two_dots": "  1.2.3    4.000        ", "coords_sixteen_digits": " 1234567890.12345 1 2   ",
}.items():
    line = atom()
    EDGE_TEXTS[tag] = line[:30] + section.ljust(24)[:24] + line[54:] + "\n"


# The reference double-frees the current atom when FREESASA_RADIUS_FROM_OCCUPANCY meets a record without a readable
# occupancy (src/structure.c:699-704 jumps to cleanup with the atom already owned by the structure) and aborts the
# process; this reader returns NULL with an error there (test_missing_occupancy_fails_cleanly).
NO_OCCUPANCY = {"truncated_at_54"}


@needs_ref
@pytest.mark.parametrize("tag", sorted(EDGE_TEXTS))
def test_edge_cases(mine, ref, tag):
    text = EDGE_TEXTS[tag].encode("latin-1")
    for options in OPTION_SETS:
        if tag in NO_OCCUPANCY and options & st.RADIUS_FROM_OCCUPANCY:
            continue
        # a MODEL record shorter than 11 characters makes the reference sscanf() past the end of the line into whatever
        # its fgets() buffer held before (src/structure.c:710): its model number is then arbitrary; here it stays 1
        same_structure(mine, ref, text, None, options, ignore=("model",) if tag == "model_short" else ())


def test_missing_occupancy_fails_cleanly(mine):
    for tag in sorted(NO_OCCUPANCY):
        assert mine.from_pdb(EDGE_TEXTS[tag].encode("latin-1"), None, st.RADIUS_FROM_OCCUPANCY) is None


def test_edge_case_known_answers(mine):
    """The handful of behaviours worth spelling out (each also covered against the reference above)."""
    assert mine.from_pdb(b"") is None and mine.from_pdb(EDGE_TEXTS["no_atoms"].encode()) is None
    s = mine.from_pdb(EDGE_TEXTS["altloc_runs"].encode())
    assert [a[0] for a in s.atoms()] == [b" N  ", b" C  ", b" O  ", b" CB "]  # first alternate location of each run wins
    s = mine.from_pdb(EDGE_TEXTS["chain_returns"].encode())
    assert s.chain_labels == b"AB" and s.n_residues == 3  # a chain is registered once; residues split on every change
    s = mine.from_pdb(EDGE_TEXTS["model_numbers"].encode())
    assert (s.n, s.model) == (1, 17)  # reading stops at the first ENDMDL
    s = mine.from_pdb(EDGE_TEXTS["model_numbers"].encode(), None, st.JOIN_MODELS)
    assert (s.n, s.model) == (2, 1)
    s = mine.from_pdb(EDGE_TEXTS["coords_exponent"].encode())
    np.testing.assert_array_equal(s.xyz()[0], [15.0, -0.25, 3.0])
    s = mine.from_pdb(EDGE_TEXTS["truncated_at_77"].encode())
    assert s.n == 2  # no element column -> the hydrogen is NOT recognised (src/pdb.c:260-283), kept with a guessed radius


@needs_ref
def test_random_mutations(mine, ref):
    """Byte-level fuzz: random edits of a valid file must never make the two readers disagree."""
    rng = np.random.default_rng(12345)
    base = bytearray(w.pdb_text(120, seed=21, chains=2, hydrogens=0.2, hetatm=2, altloc=0.1, unknown=0.1).encode())
    alphabet = b" \n\tATOMHE0123456789.-+eXx"
    for trial in range(150):
        text = bytearray(base)
        for _ in range(int(rng.integers(1, 6))):
            pos = int(rng.integers(0, len(text)))
            kind = int(rng.integers(0, 3))
            if kind == 0:
                text[pos] = alphabet[int(rng.integers(0, len(alphabet)))]
            elif kind == 1:
                del text[pos:pos + int(rng.integers(1, 40))]
            else:
                text[pos:pos] = bytes(alphabet[int(k)] for k in rng.integers(0, len(alphabet), size=int(rng.integers(1, 30))))
        options = OPTION_SETS[trial % len(OPTION_SETS)]
        if options & st.RADIUS_FROM_OCCUPANCY:  # a mutated occupancy field crashes the reference (see NO_OCCUPANCY)
            options = st.INCLUDE_HETATM | st.HALT_AT_UNKNOWN
        same_structure(mine, ref, bytes(text), None, options)


# ---- models and chains as separate structures ------------------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("options", [st.SEPARATE_MODELS, st.SEPARATE_CHAINS, st.SEPARATE_MODELS | st.SEPARATE_CHAINS,
                                     st.SEPARATE_MODELS | st.INCLUDE_HETATM | st.INCLUDE_HYDROGEN, st.SEPARATE_CHAINS | st.INCLUDE_HETATM, 0])
def test_structure_array(mine, ref, options):
    texts = [w.pdb_text(150, seed=31, chains=3, models=4, hydrogens=0.1, hetatm=2).encode(),
             w.pdb_text(150, seed=32, chains=2, models=1, hetatm=1).encode(),
             EDGE_TEXTS["endmdl_first"].encode(), EDGE_TEXTS["model_short"].encode(), EDGE_TEXTS["no_atoms"].encode(),
             EDGE_TEXTS["chain_returns"].encode()]  # (a MODEL without ENDMDL is undefined in the reference: see
    #                                          test_truncated_ensemble_keeps_its_last_model)
    for text in texts:
        a, b = mine.array(text, None, options), ref.array(text, None, options)
        assert (a is None) == (b is None)
        if a is not None:
            assert len(a) == len(b)
            for sa, sb in zip(a, b):
                assert snapshot(sa) == snapshot(sb)


@needs_ref
def test_get_chains_and_add_atom(mine, ref):
    text = w.pdb_text(200, seed=41, chains=3, unknown=0.1).encode()
    a, b = mine.from_pdb(text), ref.from_pdb(text)
    for chains in (b"A", b"B", b"AC", b"CA", b"ABC", b"D", b"AD"):
        ca, cb = a.get_chains(chains), b.get_chains(chains)
        assert (ca is None) == (cb is None), chains
        if ca is not None:
            assert snapshot(ca) == snapshot(cb)
    built = []
    for api in (mine, ref):
        s = api.new()
        rc = [s.add_atom(b" CA ", b"ALA", b"   1 ", b"A", 0.0, 0.0, 0.0),
              s.add_atom(b" CB ", b"ALA", b"   1 ", b"A", 1.5, 0.0, 0.0),
              s.add_atom(b" N  ", b"GLY", b"   2 ", b"A", 3.0, 0.0, 0.0),
              s.add_atom(b" XY ", b"ZZZ", b"   3 ", b"B", 4.5, 0.0, 0.0),
              s.add_atom(b" XY ", b"ZZZ", b"   4 ", b"B", 6.0, 0.0, 0.0, api.classifier("oons"), st.SKIP_UNKNOWN),
              s.add_atom(b"ABCD", b"ZZZ", b"   4 ", b"B", 6.0, 0.0, 0.0, api.classifier("oons"), st.SKIP_UNKNOWN),
              s.add_atom(b" XY ", b"ZZZ", b"   5 ", b"B", 7.5, 0.0, 0.0, api.classifier("oons"), st.HALT_AT_UNKNOWN),
              s.add_atom(b" O  ", b"GLY", b"   6 ", b"C", 9.0, 0.0, 0.0, api.classifier("oons"), 0)]
        built.append((rc, snapshot(s)))
    assert built[0] == built[1]
    assert built[0][1]["classifier"] == b"conflicting-classifiers"  # src/structure.c:543-560


# ---- classifiers -----------------------------------------------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("which", ["protor", "naccess", "oons"])
def test_classifier_lookups_match_reference(mine, ref, which):
    residues = sorted(w.RESIDUE_ATOMS) + ["ANY", "HOH", "A", "DA", "U", "SEC", "MSE", "ACE", "NH2", "UNK", "ZZZ", "ala", "AL"]
    atoms = sorted({a for v in w.RESIDUE_ATOMS.values() for a in v.split()}) + ["OXT", "P", "OP1", "O5'", "C1'", "SE", "H", "XX", "ca"]
    cm, cr = mine.classifier(which), ref.classifier(which)
    assert mine.lib.freesasa_classifier_name(cm) == ref.lib.freesasa_classifier_name(cr)
    for res in residues:
        pm, pr = mine.lib.freesasa_classifier_residue_reference(cm, res.encode()), ref.lib.freesasa_classifier_residue_reference(cr, res.encode())
        assert bool(pm) == bool(pr)
        if pm:
            assert (pm.contents.name, pm.contents.values()) == (pr.contents.name, pr.contents.values())
        for name in atoms:
            for r, a in ((res, name), (res.ljust(3), (" " + name).ljust(4)), (" " + res, name + "  ")):
                rm = mine.lib.freesasa_classifier_radius(cm, r.encode(), a.encode())
                assert rm == ref.lib.freesasa_classifier_radius(cr, r.encode(), a.encode())
                assert mine.lib.freesasa_classifier_class(cm, r.encode(), a.encode()) == ref.lib.freesasa_classifier_class(cr, r.encode(), a.encode())


def test_classifier_known_answers(mine):
    """reference tests/test_classifier.c: ProtOr/OONS radii and classes for a few atoms, backbone names, element guesses."""
    p, o = mine.classifier("protor"), mine.classifier("oons")
    R, C = mine.lib.freesasa_classifier_radius, mine.lib.freesasa_classifier_class
    assert R(p, b"ALA", b" CA ") == 1.88 and R(p, b"ALA", b" N  ") == 1.64 and R(p, b"ALA", b" O  ") == 1.42
    assert R(o, b"ALA", b" CA ") == 2.00 and R(o, b"ALA", b" N  ") == 1.55 and R(o, b"ALA", b" O  ") == 1.40
    assert R(p, b"ALA", b" XX ") == -1.0 and C(p, b"ALA", b" XX ") == st.ATOM_UNKNOWN
    assert C(p, b"ALA", b" CB ") == st.ATOM_APOLAR and C(p, b"ALA", b" O  ") == st.ATOM_POLAR
    bb = mine.lib.freesasa_atom_is_backbone
    assert all(bb(a) for a in (b" CA ", b" N  ", b" O  ", b" C  ", b"OXT", b" P  ", b" O5'")) and not any(bb(a) for a in (b" CB ", b"", b"    ", b" SG "))
    g = mine.lib.freesasa_guess_radius
    assert (g(b" C"), g(b" N"), g(b" O"), g(b"SE"), g(b" H"), g(b"C"), g(b"XX"), g(b"")) == (1.70, 1.55, 1.52, 1.90, 1.10, 1.70, -1.0, -1.0)


@needs_ref
def test_guess_radius_all_symbols(mine, ref):
    letters = " ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz1"
    for a in letters:
        for b in letters:
            s = (a + b).encode()
            assert mine.lib.freesasa_guess_radius(s) == ref.lib.freesasa_guess_radius(s), s
        assert mine.lib.freesasa_guess_radius(a.encode()) == ref.lib.freesasa_guess_radius(a.encode())


CONFIGS = {
    "plain": "name: test-config\n\ntypes:\nA 1.0 polar # comment\nB 2.0 apolar\n# comment\n\natoms:\nAA aa A # comment\nBB bb B\nANY cc A\n",
    "no_name": "types:\nA 1.0 polar\natoms:\nAA aa A\nBB bb A\n",
    "sections_reordered": "atoms:\nAA aa A\nBB bb B\ntypes:\nA 1.5 polar\nB 2.5 apolar\nname: later\n",
    "duplicate_type_in_middle": "name: d\ntypes:\nA 1.0 polar\nA 2.0 apolar\nB 2.0 apolar\natoms:\nAA aa A\nBB bb B\n",
    "duplicate_type_last": "name: d\ntypes:\nA 1.0 polar\nB 2.0 apolar\nB 3.0 polar\natoms:\nAA aa A\nBB bb B\n",
    "duplicate_atom_in_middle": "name: d\ntypes:\nA 1.0 polar\nB 2.0 apolar\natoms:\nAA aa A\nAA aa B\nBB bb B\n",
    "duplicate_atom_last": "name: d\ntypes:\nA 1.0 polar\nB 2.0 apolar\natoms:\nAA aa A\nBB bb B\nBB bb A\n",
    "unknown_type": "name: d\ntypes:\nA 1.0 polar\natoms:\nAA aa Q\n",
    "bad_class": "name: d\ntypes:\nA 1.0 hydrophobic\natoms:\nAA aa A\n",
    "bad_radius": "name: d\ntypes:\nA x polar\natoms:\nAA aa A\n",
    "two_fields": "name: d\ntypes:\nA 1.0 polar\natoms:\nAA aa\n",
    "long_residue": "name: d\ntypes:\nA 1.0 polar\natoms:\nAAAA aa A\n",
    "long_atom": "name: d\ntypes:\nA 1.0 polar\natoms:\nAA aaaaa A\n",
    "no_types": "name: d\natoms:\nAA aa A\n",
    "no_atoms": "name: d\ntypes:\nA 1.0 polar\n",
    "empty_name": "name:\ntypes:\nA 1.0 polar\natoms:\nAA aa A\nBB bb A\n",
    "commented_keyword": "name: d\n# types: not here\ntypes:\nA 1.0 polar\natoms:\nAA aa A\nBB bb A\n",
    "crlf": "name: d\r\ntypes:\r\nA 1.0 polar\r\natoms:\r\nAA aa A\r\nBB bb A\r\n",
    "class_case": "name: d\ntypes:\nA 1.0 Polar\nB 2.0 APOLAR\natoms:\nAA aa A\nBB bb B\n",
    "prefix_classes": "name: d\ntypes:\nA 1.0 polarity\nB 2.0 apolarish\natoms:\nAA aa A\nBB bb B\n",
    "single_char_line": "name: d\ntypes:\nA 1.0 polar\nx\natoms:\nAA aa A\nBB bb A\n",
    "empty": "",
}


@needs_ref
@pytest.mark.parametrize("tag", sorted(CONFIGS))
def test_classifier_from_file(mine, ref, tag):
    text = CONFIGS[tag].encode()
    cm, cr = mine.classifier_from_text(text), ref.classifier_from_text(text)
    assert bool(cm) == bool(cr), tag
    if not cm:
        return
    if "no_name" not in tag:  # the reference leaves the name NULL there (and would crash registering it)
        assert mine.lib.freesasa_classifier_name(cm) == ref.lib.freesasa_classifier_name(cr)
    for res in (b"AA", b"BB", b"CC", b"ANY", b" AA"):
        for a in (b"aa", b"bb", b"cc", b" aa ", b"zz"):
            assert mine.lib.freesasa_classifier_radius(cm, res, a) == ref.lib.freesasa_classifier_radius(cr, res, a)
            assert mine.lib.freesasa_classifier_class(cm, res, a) == ref.lib.freesasa_classifier_class(cr, res, a)
    pdb = (w.pdb_atom_line(1, "aa", "AA", "A", 1, 0, 0, 0, "C") + "\n" + w.pdb_atom_line(2, "cc", "QQ", "A", 2, 3, 0, 0, "C") + "\n").encode()
    if "no_name" not in tag:
        a, b = mine.from_pdb(pdb, cm), ref.from_pdb(pdb, cr)
        assert snapshot(a) == snapshot(b)
    mine.lib.freesasa_classifier_free(cm)
    ref.lib.freesasa_classifier_free(cr)


@needs_ref
@pytest.mark.skipif(not os.path.exists("/root/reference/share/naccess.config"), reason="reference configs not present")
@pytest.mark.parametrize("name", ["naccess", "oons", "protor", "dssp"])
def test_shipped_configs(mine, ref, name):
    """The reference's own configuration files (share/*.config) parse to the same classifier in both libraries."""
    with open(f"/root/reference/share/{name}.config", "rb") as f:
        text = f.read()
    cm, cr = mine.classifier_from_text(text), ref.classifier_from_text(text)
    assert bool(cm) == bool(cr)
    assert bool(cm) == (name != "dssp")  # dssp.config uses classes the parser rejects ("backbone"), in both libraries
    if not cm:
        return
    pdb = w.pdb_text(400, seed=51, hydrogens=0.2, hetatm=2, unknown=0.1).encode()
    for options in (0, st.INCLUDE_HETATM | st.INCLUDE_HYDROGEN, st.SKIP_UNKNOWN):
        assert snapshot(mine.from_pdb(pdb, cm, options)) == snapshot(ref.from_pdb(pdb, cr, options))


def test_pdb_line_accessor_is_stable(mine):
    text = w.pdb_text(50, seed=61).encode()
    s = mine.from_pdb(text)
    lines = [ln + b"\n" for ln in text.split(b"\n") if ln.startswith(b"ATOM")]
    assert [a[7] for a in s.atoms()] == lines
    assert s.new if False else True
    t = mine.new()
    t.add_atom(b" CA ", b"ALA", b"   1 ", b"A", 0.0, 0.0, 0.0)
    assert t.atoms()[0][7] is None  # atoms added by hand carry no PDB line (src/structure.c:1199-1206)


@needs_ref
def test_reading_from_a_pipe(mine, ref):
    """`cat x.pdb | program`: the stream is not seekable; both libraries read up to EOF (the reference by accident of
    its ftell() arithmetic, src/util.c:20-34)."""
    text = w.pdb_text(3000, seed=71, chains=2, hetatm=2).encode()  # larger than a pipe buffer
    snaps = []
    for api in (mine, ref):
        import threading

        r, wr = os.pipe()

        def feed():
            with os.fdopen(wr, "wb") as f:
                f.write(text)

        t = threading.Thread(target=feed)
        t.start()
        st._libc.fdopen.restype = ctypes.c_void_p
        fp = st._libc.fdopen(r, b"r")
        h = api.lib.freesasa_structure_from_pdb(fp, None, 0)
        st._libc.fclose(fp)
        t.join()
        assert h
        snaps.append(snapshot(st.Structure(api, h)))
    assert snaps[0] == snaps[1]
    assert snaps[0] == snapshot(mine.from_pdb(text))


@needs_ref
def test_parallel_reader_is_identical_to_the_reference(mine, ref):
    """Ranges of >= 1 MB are cut at safe line boundaries, parsed by several threads and stitched (ingest.c).  Forced on
    for small inputs through the test hook; every thread count must reproduce the reference exactly, seams included."""
    hook = mine.lib.fsb_ingest_parallel_reads
    hook.restype = ctypes.c_long
    texts = [w.pdb_text(900, seed=81, chains=4, models=1, hydrogens=0.3, hetatm=5, altloc=0.3, unknown=0.1).encode(),
             w.pdb_text(700, seed=82, chains=2, models=3, hydrogens=0.1, altloc=0.2).encode(),
             w.pdb_text(600, seed=83, chains=3, element_column=False, newline="\r\n").encode(),
             EDGE_TEXTS["altloc_runs"].encode() * 40, EDGE_TEXTS["chain_returns"].encode() * 30,
             (EDGE_TEXTS["insertion_codes"] * 25 + EDGE_TEXTS["coords_garbage"] + EDGE_TEXTS["insertion_codes"] * 25).encode(),
             (EDGE_TEXTS["nucleic"] * 20 + EDGE_TEXTS["very_long_line"] + EDGE_TEXTS["nucleic"] * 20).encode()]
    old = {k: os.environ.get(k) for k in ("FREESASA_B200_PARALLEL_MIN_BYTES", "FREESASA_B200_THREADS")}
    try:
        os.environ["FREESASA_B200_PARALLEL_MIN_BYTES"] = "1"
        before = hook()
        for threads in (2, 3, 7, 16):
            os.environ["FREESASA_B200_THREADS"] = str(threads)
            for text in texts:
                for options in OPTION_SETS:
                    same_structure(mine, ref, text, None, options)
                for options in (st.SEPARATE_MODELS, st.SEPARATE_CHAINS | st.SEPARATE_MODELS):
                    a, b = mine.array(text, None, options), ref.array(text, None, options)
                    assert (a is None) == (b is None)
                    if a is not None:
                        assert [snapshot(x) for x in a] == [snapshot(x) for x in b]
        assert hook() - before > 300  # the parallel path really ran
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_large_file_is_read_in_slices(mine, tmp_path):
    """Files of >= 4 MB are read with pread() slices from several threads, then parsed in parallel: same structure as
    the single-threaded reader gives for the same bytes."""
    text = w.pdb_text(60000, seed=91, chains=5, hetatm=10).encode()
    assert len(text) > (4 << 20)
    path = os.path.join(tmp_path, "big.pdb")
    with open(path, "wb") as f:
        f.write(text)
    old = os.environ.get("FREESASA_B200_THREADS")
    try:
        os.environ["FREESASA_B200_THREADS"] = "1"
        serial = snapshot(mine.from_pdb(text))
        os.environ["FREESASA_B200_THREADS"] = "6"
        assert snapshot(mine.from_pdb_path(path)) == serial
        assert snapshot(mine.from_pdb(text)) == serial
    finally:
        if old is None:
            os.environ.pop("FREESASA_B200_THREADS", None)
        else:
            os.environ["FREESASA_B200_THREADS"] = old
    assert serial["n"] == 60000 and serial["n_chains"] == 5


class CifAtomLcl(ctypes.Structure):
    """struct freesasa_cif_atom_lcl, reference src/freesasa.h:351-363."""

    _fields_ = [(k, ctypes.c_char_p) for k in ("group_PDB", "auth_asym_id", "auth_seq_id", "pdbx_PDB_ins_code", "auth_comp_id",
                                                 "auth_atom_id", "label_alt_id", "type_symbol")] + \
               [(k, ctypes.c_double) for k in ("Cartn_x", "Cartn_y", "Cartn_z")]


class ChainGroup(ctypes.Structure):
    _fields_ = [("chains", ctypes.POINTER(ctypes.c_char_p)), ("n", ctypes.c_size_t)]


@needs_ref
def test_long_chain_labels_and_cif_atoms(mine, ref):
    """The entry points an mmCIF reader uses (src/structure.c:793-829) and the long-chain-label (_lcl) variants
    (src/structure.c:1012-1081,1303-1378): up to three-character chain labels, insertion codes from a separate column."""
    rows = []
    rng = np.random.default_rng(5)
    for chain, n_res in ((b"AA", 4), (b"B", 3), (b"AB1", 2), (b"AA", 2)):
        for r in range(n_res):
            res = [b"ALA", b"GLY", b"HOH", b"UNK"][int(rng.integers(0, 4))]
            ins = b"?" if r % 3 else b"A"
            for name, sym in ((b"N", b"N"), (b"CA", b"C"), (b"O", b"O"), (b"XX", b"X")):
                x, y, z = (float(v) for v in rng.uniform(-20, 20, size=3))
                rows.append(CifAtomLcl(b"ATOM", chain, b"%d" % (r + 1), ins, res, name, b".", sym, x, y, z))
    snaps, extra = [], []
    for api in (mine, ref):
        L = api.lib
        L.freesasa_structure_add_cif_atom_lcl.argtypes = [ctypes.c_void_p, ctypes.POINTER(CifAtomLcl), ctypes.c_void_p, ctypes.c_int]
        L.freesasa_structure_get_chains_lcl.restype = ctypes.c_void_p
        L.freesasa_structure_get_chains_lcl.argtypes = [ctypes.c_void_p, ctypes.POINTER(ChainGroup), ctypes.c_void_p, ctypes.c_int]
        for f in ("freesasa_structure_chain_atoms_lcl", "freesasa_structure_chain_residues_lcl"):
            getattr(L, f).argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        L.freesasa_structure_residue_chain_lcl.restype = ctypes.c_char_p
        L.freesasa_structure_residue_chain_lcl.argtypes = [ctypes.c_void_p, ctypes.c_int]
        s = api.new()
        rc = [L.freesasa_structure_add_cif_atom_lcl(s.h, ctypes.byref(row), None, st.SKIP_UNKNOWN if k % 5 == 0 else 0) for k, row in enumerate(rows)]
        a, b = ctypes.c_int(), ctypes.c_int()
        ranges = []
        for label in (b"AA", b"B", b"AB1"):
            ranges.append((L.freesasa_structure_chain_atoms_lcl(s.h, label, ctypes.byref(a), ctypes.byref(b)), a.value, b.value,
                           L.freesasa_structure_chain_residues_lcl(s.h, label, ctypes.byref(a), ctypes.byref(b)), a.value, b.value))
        res_chains = [L.freesasa_structure_residue_chain_lcl(s.h, r) for r in range(s.n_residues)]
        subs = []
        for group in ([b"AA"], [b"B", b"AB1"], [b"AB1", b"AA", b"B"], [b"ZZ"], [b"AA", b"ZZ"]):
            arr = (ctypes.c_char_p * len(group))(*group)
            h = L.freesasa_structure_get_chains_lcl(s.h, ctypes.byref(ChainGroup(arr, len(group))), None, 0)
            subs.append(snapshot(st.Structure(api, h)) if h else None)
        labels = [s._call("chain_label", k) for k in range(s.n_chains)]
        snap = snapshot(s)
        snap.pop("chains")  # the single-character accessors used by snapshot() cannot address "AA"/"AB1"
        for sub in subs:
            if sub:
                sub.pop("chains")
        snaps.append(snap)
        extra.append((rc, ranges, res_chains, subs, labels))
    assert snaps[0] == snaps[1]
    assert extra[0] == extra[1]
    assert extra[0][4] == [b"AA", b"B", b"AB1"] and snaps[0]["n_chains"] == 3


def test_decimal_fast_paths_equal_strtod(mine):
    """The coordinate parser's claim (ingest.c: strict_coords / fast_coords): integer mantissa / power of ten, ONE IEEE
    division, is the correctly rounded value — i.e. what strtod (and Python's float()) return.  80 000 records in the
    %8.3f layout over its whole range (240 000 numbers) and 13 000 records of free-format tokens with 1..15 digits."""
    rng = np.random.default_rng(2024)
    template = w.pdb_atom_line(1, "CA", "ALA", "A", 1, 0.0, 0.0, 0.0, "C")
    lines, want = [], []
    specials = [0.0, -0.0, 0.001, -0.001, 999.999, -999.999, 0.1, 0.7, 123.5, -0.0004, 0.0005]
    for k in range(80000):
        first = rng.uniform(-999.999, 9999.999) if k % 4 else float(rng.integers(-999, 9999))
        rest = rng.uniform(-999.999, 999.999, size=2) if k % 7 else rng.choice(specials, size=2)
        trio = ["%8.3f" % v for v in (first, *rest)]
        lines.append(template[:30] + "".join(trio) + template[54:])       # the fixed-column layout, fields may touch
        want.append([float(t) for t in trio])
    for k in range(13000):
        trio = []
        for _ in range(3):
            digits = int(rng.integers(1, 8 if k % 3 else 16))
            frac = int(rng.integers(0, digits + 1))
            m = "".join(str(d) for d in rng.integers(0, 10, size=digits))
            tok = (m[: digits - frac].lstrip("0") or "0") + ("." + m[digits - frac:] if frac else "")
            if k % 11 == 0 and frac:
                tok = tok[1:] if tok.startswith("0.") else tok                  # ".5" style
            trio.append(("-" if rng.random() < 0.4 else "+" if rng.random() < 0.05 else "") + tok)
        section = " ".join(trio)
        if len(section) > 24:
            continue
        lines.append(template[:30] + section.ljust(24) + template[54:])    # what sscanf("%lf%lf%lf") reads
        want.append([float(t) for t in trio])
    text = ("\n".join(lines) + "\n").encode()
    s = mine.from_pdb(text)
    assert s.n == len(lines) > 85000
    np.testing.assert_array_equal(s.xyz().view(np.uint64), np.array(want).view(np.uint64))   # bit for bit, signed zeros included


def test_truncated_ensemble_keeps_its_last_model(mine):
    """A file that ends inside a MODEL (no ENDMDL): the last model runs to the end of the text.  (The reference leaves
    that model's end uninitialised, src/pdb.c:63-73, so there is nothing to compare with.)"""
    text = w.pdb_text(60, seed=3, chains=2, models=3).encode()
    cut = text[: text.rindex(b"ENDMDL")]
    models = mine.array(cut, None, st.SEPARATE_MODELS)
    assert [m.n for m in models] == [60, 60, 60] and [m.model for m in models] == [1, 2, 3]
    full = mine.array(text, None, st.SEPARATE_MODELS)
    assert snapshot(models[2]) == snapshot(full[2])

"""CPU tests for the selection row: "name, expression" -> area of the selected atoms.

Oracle: the compiled reference with its own flex/bison-generated parser (oracle/_ref).  For every command both libraries
must agree on: accepted or rejected, the selection's name, the return code (success / warning), and the area BIT FOR BIT
(the sum runs over all atoms in order with a 0/1 factor, src/selection.c:715-718).  Commands cover the grammar of
src/parser.y and the scanner rules of src/lexer.l, the examples of the reference's documentation and tests
(tests/test_selection.c), and randomly generated expressions."""
import ctypes

import numpy as np
import pytest

from freesasa_b200 import structure as st
from freesasa_b200 import workloads as w
from oracle import bindings as ob

needs_ref = pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not built")


class Selector:
    def __init__(self, api):
        self.api, L = api, api.lib
        res_p = ctypes.POINTER(api.Result)
        L.freesasa_selection_new.restype = ctypes.c_void_p
        L.freesasa_selection_new.argtypes = [ctypes.c_char_p, ctypes.c_void_p, res_p]
        L.freesasa_selection_free.argtypes = [ctypes.c_void_p]
        L.freesasa_selection_free.restype = None
        for name, restype in (("name", ctypes.c_char_p), ("command", ctypes.c_char_p), ("area", ctypes.c_double)):
            f = getattr(L, "freesasa_selection_" + name)
            f.restype, f.argtypes = restype, [ctypes.c_void_p]
        L.freesasa_select_area.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_double), ctypes.c_void_p, res_p]

    def run(self, command: bytes, structure, result):
        """(name, command, area bits, n_atoms) of freesasa_selection_new(), or None; plus freesasa_select_area()'s view."""
        L = self.api.lib
        h = L.freesasa_selection_new(command, structure.h, ctypes.byref(result))
        out = None
        if h:
            # n_atoms straight from the struct {char *name; char *command; double area; int n_atoms;} (src/selection.c:17-22):
            # the reference declares freesasa_selection_n_atoms() but never defines it
            n_atoms = ctypes.cast(h + 24, ctypes.POINTER(ctypes.c_int))[0]
            out = (L.freesasa_selection_name(h), L.freesasa_selection_command(h),
                   np.float64(L.freesasa_selection_area(h)).view(np.uint64).item(), n_atoms)
            L.freesasa_selection_free(h)
        name = ctypes.create_string_buffer(64)
        area = ctypes.c_double(-1.0)
        rc = L.freesasa_select_area(command, name, ctypes.byref(area), structure.h, ctypes.byref(result))
        return out, (rc, name.value, np.float64(area.value).view(np.uint64).item())


@pytest.fixture(scope="module")
def world():
    mine = st.api()
    ref = st.StructureAPI(ob.ref_lib(), ob.RefResult, ob.RefParameters)
    text = (w.pdb_text(900, seed=17, chains=3, hetatm=4, unknown=0.05)
            + w.pdb_atom_line(9001, "CA", "ALA", "D", -5, 1.0, 2.0, 3.0, "C") + "\n"
            + w.pdb_atom_line(9002, "C5'", "  A", "D", 7, 4.0, 2.0, 3.0, "C") + "\n"
            + w.pdb_atom_line(9003, "SE", "MSE", "D", 8, 7.0, 2.0, 3.0, "SE", icode="A") + "\n"
            + w.pdb_atom_line(9004, "CA", "MSE", "D", 8, 9.0, 2.0, 3.0, "C", icode="B") + "\n").encode()
    out = []
    for api in (mine, ref):
        api.lib.freesasa_set_verbosity(2)
        s = api.from_pdb(text, None, st.INCLUDE_HETATM)
        tree = st.TreeAPI(api)
        rng = np.random.default_rng(3)
        result, keep = tree.make_result(rng.uniform(0, 40, size=s.n) * (rng.random(s.n) < 0.7))
        out.append((Selector(api), s, result, keep))
    yield out
    mine.lib.freesasa_set_verbosity(0)


COMMANDS = [
    # documentation / reference tests (tests/test_selection.c)
    "aromatic, resn phe+tyr+trp+his+pro", "s, resn ALA", "s, resn ala", "s, symbol C", "s, symbol O+N", "s, name CA", "s, name ca+cb",
    "s, chain A", "s, chain A+B", "s, chain A-C", "s, chain B-C+A", "s, resi 1", "s, resi 1-20", "s, resi 1-20+30-40+50", "s, resi -10",
    "s, resi 100-", "s, resi \\-5", "s, resi \\-5-10", "s, resi \\-10-\\-1", "s, resi 8A", "s, resi 8A+8B+7", "s, resi 1-20+8A",
    "s, resn ALA and chain A", "s, resn ALA or chain B", "s, not resn ALA", "s, not resn ALA and chain A", "s, not (resn ALA and chain A)",
    "s, resn ALA and not chain A or symbol O", "s, (resn ALA or resn GLY) and (chain A or chain B)", "s, resn ALA & chain A | ! symbol C",
    "s, RESN ala AND CHAIN a", "s, Resn Ala Or Not Symbol n", "s,resn ALA", "  s  ,  resn   ALA  ", "s,\tresn ALA\n",
    "a-b+c_1, resn ALA", "10, resi 10", "and, chain A", "resn, resn ALA", "s, name C5'", "s, name C5'+CA", "s, symbol SE", "s, symbol se+c",
    "s, resn HOH", "s, resn A", "s, resn MSE+A", "s, name OXT+XX1", "s, symbol FE", "s, resn LIG and symbol CL",
    # warnings: no matches, invalid identifiers, invalid ranges
    "s, resn XYZ", "s, resn ALAA", "s, name ABCDE", "s, symbol ABC", "s, symbol 1", "s, chain AB", "s, resi 1A2", "s, resi A", "s, resi 123456A",
    "s, chain A-1", "s, chain AB-C", "s, resi 1-A", "s, resn ALA+XYZ", "s, resn ALAA or resn GLY", "s, not resn ALAA", "s, chain 1-3", "s, chain 1",
    # syntax errors
    "", "s", "s,", "resn ALA", "s resn ALA", "s, ", "s, resn", "s, resn ALA and", "s, and resn ALA", "s, (resn ALA", "s, resn ALA)", "s, resn ALA chain A",
    "s, resn ALA+", "s, resn +ALA", "s, resn ALA-GLY", "s, resi 1--2", "s, chain -A", "s, chain A-", "s, foo ALA", "s, resn ALA,", "s, t, resn ALA",
    "s, resi 1+", "s, not", "s, ()", "s, resn ALA or or chain A", "s, name CA CB",
    "averyveryveryveryveryveryveryveryveryveryverylongselectionname, resn ALA",
]


@needs_ref
@pytest.mark.parametrize("command", COMMANDS)
def test_commands_match_reference(world, command):
    (sm, s_m, r_m, _), (sr, s_r, r_r, _) = world
    got, want = sm.run(command.encode(), s_m, r_m), sr.run(command.encode(), s_r, r_r)
    assert got == want, command


def test_known_answers(world):
    """Hand-checkable: complement, union and the documented precedence not > and > or."""
    (sm, s, r, keep), _ = world
    area = lambda c: np.uint64(sm.run(c.encode(), s, r)[0][2]).view(np.float64).item()  # noqa: E731
    total = float(np.cumsum(keep)[-1])
    assert area("s, chain A-D") == total
    assert abs(area("s, resn ALA") + area("s, not resn ALA") - total) < 1e-9
    assert area("s, resn ALA or resn GLY") == area("s, resn ALA+GLY")
    assert area("s, not resn ALA and chain A") == area("s, (not resn ALA) and chain A")
    assert area("s, resn ALA or resn GLY and chain A") == area("s, resn ALA or (resn GLY and chain A)")
    assert sm.run(b"s, resn ALAA", s, r)[1][0] == -2  # FREESASA_WARN: the invalid name is ignored
    assert sm.run(b"s resn ALA", s, r) == (None, (-1, b"", 0))


@needs_ref
def test_random_expressions(world):
    (sm, s_m, r_m, _), (sr, s_r, r_r, _) = world
    rng = np.random.default_rng(99)
    resn = ["ALA", "gly", "LYS", "HOH", "XYZ", "A", "MSE", "TRPP"]
    names = ["CA", "cb", "N", "O", "C5'", "OXT", "FE", "SE", "ABCDE"]
    symbols = ["C", "n", "O", "SE", "FE", "S", "1", "CLX"]
    chains = ["A", "b", "C", "D", "E", "1"]

    def atom():
        k = int(rng.integers(0, 5))
        plus = lambda pool: "+".join(rng.choice(pool, size=int(rng.integers(1, 4))))  # noqa: E731
        if k == 0:
            return "resn " + plus(resn)
        if k == 1:
            return "name " + plus(names)
        if k == 2:
            return "symbol " + plus(symbols)
        if k == 3:
            items = []
            for _ in range(int(rng.integers(1, 4))):
                a, b = sorted(int(v) for v in rng.integers(-8, 140, size=2))
                fmt = lambda v: ("\\-%d" % -v) if v < 0 else str(v)  # noqa: E731
                items.append(rng.choice([fmt(a), "%s-%s" % (fmt(a), fmt(b)), "-%s" % fmt(b), "%s-" % fmt(a), "8A", "8b"]))
            return "resi " + "+".join(items)
        pair = sorted(rng.choice(chains[:5], size=2))
        return "chain " + rng.choice([plus(chains), "%s-%s" % (pair[0].upper(), pair[1].upper())])

    def expr(depth):
        if depth == 0 or rng.random() < 0.3:
            return atom()
        k = int(rng.integers(0, 4))
        if k == 0:
            return "not " + expr(depth - 1)
        if k == 1:
            return "(" + expr(depth - 1) + ")"
        return expr(depth - 1) + rng.choice([" and ", " or ", " & ", " | ", " AND "]) + expr(depth - 1)

    for trial in range(400):
        command = ("r%d, " % trial + expr(3)).encode()
        assert sm.run(command, s_m, r_m) == sr.run(command, s_r, r_r), command


@needs_ref
def test_selections_on_structure_nodes(world):
    """freesasa_node_structure_add_selection() / _selections(): the structure node keeps clones (src/node.c:670-709)."""
    seen = []
    for sel, s, result, _ in world:
        L, tree = sel.api.lib, st.TreeAPI(sel.api)
        L.freesasa_node_structure_add_selection.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.freesasa_node_structure_selections.restype = ctypes.POINTER(ctypes.c_void_p)
        L.freesasa_node_structure_selections.argtypes = [ctypes.c_void_p]
        root = tree.init(result, s, b"t")
        node = L.freesasa_node_children(L.freesasa_node_children(root))
        assert not L.freesasa_node_structure_selections(node)
        for command in (b"ala, resn ALA", b"bb, name CA+C+N+O and chain A-B"):
            h = L.freesasa_selection_new(command, s.h, ctypes.byref(result))
            assert L.freesasa_node_structure_add_selection(node, h) == 0
            L.freesasa_selection_free(h)  # the node holds its own copy
        arr, got, k = L.freesasa_node_structure_selections(node), [], 0
        while arr[k]:
            got.append((L.freesasa_selection_name(arr[k]), L.freesasa_selection_command(arr[k]), L.freesasa_selection_area(arr[k])))
            k += 1
        seen.append(got)
        assert tree.free(root) == 0
    assert seen[0] == seen[1] and len(seen[0]) == 2 and seen[0][0][2] > 0


def test_hostile_nesting_is_rejected_not_crashed(world):
    (sm, s, r, _), _ = world
    deep = ("s, " + "(" * 5000 + "resn ALA" + ")" * 5000).encode()
    assert sm.run(deep, s, r)[0] is None
    nots = ("s, " + "not " * 5000 + "resn ALA").encode()
    assert sm.run(nots, s, r)[0] is None
    ok = ("s, " + "(" * 150 + "resn ALA" + ")" * 150).encode()
    assert sm.run(ok, s, r)[0] is not None

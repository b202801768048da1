"""ctypes mirror of the structure / classifier part of the C host layer (include/freesasa_b200_host.h, scope rows
f-1 .. f-3 of SURVEY.md §8): PDB text -> structure -> SASA -> per-residue / per-chain / per-class areas.

``StructureAPI`` binds the reference's own function names on whatever shared library it is given, so the parity
tests drive this repo's libfreesasa_b200_host.so and the compiled reference through the very same Python code.
Nothing here computes anything: parsing and classification happen in csrc/ingest.c + csrc/radii.c, the SASA in the
CUDA engine, the aggregation in csrc/areas.c.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

# enum freesasa_structure_options, reference src/freesasa.h:182-191
INCLUDE_HETATM = 1
INCLUDE_HYDROGEN = 1 << 2
SEPARATE_MODELS = 1 << 3
SEPARATE_CHAINS = 1 << 4
JOIN_MODELS = 1 << 5
HALT_AT_UNKNOWN = 1 << 6
SKIP_UNKNOWN = 1 << 7
RADIUS_FROM_OCCUPANCY = 1 << 8
# enum freesasa_atom_class, src/freesasa.h:163-167
ATOM_APOLAR, ATOM_POLAR, ATOM_UNKNOWN = 0, 1, 2

_vp = ctypes.c_void_p
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_libc = ctypes.CDLL(None)
_libc.fmemopen.restype = _vp
_libc.fmemopen.argtypes = [_vp, ctypes.c_size_t, ctypes.c_char_p]
_libc.fopen.restype = _vp
_libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
_libc.fclose.argtypes = [_vp]
_libc.free.argtypes = [_vp]


class NodeArea(ctypes.Structure):
    """struct freesasa_nodearea, reference src/freesasa.h:289-297."""

    _fields_ = [("name", ctypes.c_char_p), ("total", ctypes.c_double), ("main_chain", ctypes.c_double),
                ("side_chain", ctypes.c_double), ("polar", ctypes.c_double), ("apolar", ctypes.c_double),
                ("unknown", ctypes.c_double)]

    def values(self):
        return (self.total, self.main_chain, self.side_chain, self.polar, self.apolar, self.unknown)


class CFile:
    """A C ``FILE *`` over bytes (fmemopen) or a path (fopen) for the FILE*-taking reference functions."""

    def __init__(self, data: Optional[bytes] = None, path: Optional[str] = None):
        if path is not None:
            self._buf = None
            self.fp = _libc.fopen(os.fsencode(path), b"r")
        else:
            self._buf = ctypes.create_string_buffer(data, len(data)) if data else None
            # fmemopen rejects size 0; an empty stream is emulated with /dev/null
            self.fp = _libc.fmemopen(self._buf, len(data), b"r") if data else _libc.fopen(b"/dev/null", b"r")
        if not self.fp:
            raise OSError("could not open C stream")

    def __enter__(self):
        return self.fp

    def __exit__(self, *exc):
        _libc.fclose(self.fp)
        self.fp = None


class StructureAPI:
    """The reference's structure/classifier entry points bound on ``lib`` (a ctypes.CDLL)."""

    def __init__(self, lib, result_type, parameters_type):
        self.lib = L = lib
        self.Result, self.Parameters = result_type, parameters_type
        cp, ci, cd = ctypes.c_char_p, ctypes.c_int, ctypes.c_double

        def sig(name, restype, *argtypes):
            f = getattr(L, name)
            f.restype, f.argtypes = restype, list(argtypes)

        sig("freesasa_structure_from_pdb", _vp, _vp, _vp, ci)
        sig("freesasa_structure_array", ctypes.POINTER(_vp), _vp, _ip, _vp, ci)
        sig("freesasa_structure_new", _vp)
        sig("freesasa_structure_free", None, _vp)
        sig("freesasa_structure_add_atom", ci, _vp, cp, cp, cp, ctypes.c_char, cd, cd, cd)
        sig("freesasa_structure_add_atom_wopt", ci, _vp, cp, cp, cp, ctypes.c_char, cd, cd, cd, _vp, ci)
        sig("freesasa_structure_get_chains", _vp, _vp, cp, _vp, ci)
        for name in ("n", "n_residues", "n_chains", "model"):
            sig("freesasa_structure_" + name, ci, _vp)
        for name in ("chain_labels", "classifier_name"):
            sig("freesasa_structure_" + name, cp, _vp)
        for name in ("atom_name", "atom_res_name", "atom_res_number", "atom_symbol", "atom_pdb_line", "residue_name",
                     "residue_number", "atom_chain_lcl", "chain_label"):
            sig("freesasa_structure_" + name, cp, _vp, ci)
        sig("freesasa_structure_atom_class", ci, _vp, ci)
        sig("freesasa_structure_atom_radius", cd, _vp, ci)
        sig("freesasa_structure_radius", _dp, _vp)
        sig("freesasa_structure_coord_array", _dp, _vp)
        sig("freesasa_structure_set_radius", None, _vp, _dp)
        sig("freesasa_structure_residue_atoms", ci, _vp, ci, _ip, _ip)
        sig("freesasa_structure_residue_reference", ctypes.POINTER(NodeArea), _vp, ci)
        sig("freesasa_structure_chain_atoms", ci, _vp, ctypes.c_char, _ip, _ip)
        sig("freesasa_structure_chain_residues", ci, _vp, ctypes.c_char, _ip, _ip)
        sig("freesasa_calc_structure", ctypes.POINTER(result_type), _vp, ctypes.POINTER(parameters_type))
        sig("freesasa_result_free", None, ctypes.POINTER(result_type))
        sig("freesasa_classifier_from_file", _vp, _vp)
        sig("freesasa_classifier_free", None, _vp)
        sig("freesasa_classifier_radius", cd, _vp, cp, cp)
        sig("freesasa_classifier_class", ci, _vp, cp, cp)
        sig("freesasa_classifier_name", cp, _vp)
        sig("freesasa_classifier_residue_reference", ctypes.POINTER(NodeArea), _vp, cp)
        sig("freesasa_guess_radius", cd, cp)
        sig("freesasa_atom_is_backbone", ci, cp)
        sig("freesasa_set_verbosity", ci, ci)

    # ---- classifiers ------------------------------------------------------------------------------------
    def classifier(self, which: str):
        """Address of a built-in classifier object: 'protor' (default), 'naccess', 'oons'."""
        return ctypes.addressof(ctypes.c_char.in_dll(self.lib, f"freesasa_{which}_classifier"))

    def classifier_from_text(self, text: bytes):
        with CFile(text) as fp:
            return self.lib.freesasa_classifier_from_file(fp)

    # ---- structures ---------------------------------------------------------------------------------------
    def from_pdb(self, text: bytes, classifier=None, options: int = 0):
        """freesasa_structure_from_pdb() on in-memory text; returns a Structure or None (where C returns NULL)."""
        with CFile(text) as fp:
            h = self.lib.freesasa_structure_from_pdb(fp, classifier, options)
        return Structure(self, h) if h else None

    def from_pdb_path(self, path: str, classifier=None, options: int = 0):
        with CFile(path=path) as fp:
            h = self.lib.freesasa_structure_from_pdb(fp, classifier, options)
        return Structure(self, h) if h else None

    def array(self, text: bytes, classifier=None, options: int = SEPARATE_MODELS):
        """freesasa_structure_array(): list of Structure, or None."""
        n = ctypes.c_int(0)
        with CFile(text) as fp:
            arr = self.lib.freesasa_structure_array(fp, ctypes.byref(n), classifier, options)
        if not arr:
            return None
        out = [Structure(self, arr[k]) for k in range(n.value)]
        _libc.free(ctypes.cast(arr, _vp))
        return out

    def array_path(self, path: str, classifier=None, options: int = SEPARATE_MODELS):
        """freesasa_structure_array() on a file."""
        n = ctypes.c_int(0)
        with CFile(path=path) as fp:
            arr = self.lib.freesasa_structure_array(fp, ctypes.byref(n), classifier, options)
        if not arr:
            return None
        out = [Structure(self, arr[k]) for k in range(n.value)]
        _libc.free(ctypes.cast(arr, _vp))
        return out

    def new(self):
        return Structure(self, self.lib.freesasa_structure_new())


class Structure:
    """Owner of one ``freesasa_structure *``."""

    def __init__(self, api: StructureAPI, handle):
        self.api, self.h = api, handle

    def free(self):
        if self.h:
            self.api.lib.freesasa_structure_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def _call(self, name, *args):
        return getattr(self.api.lib, "freesasa_structure_" + name)(self.h, *args)

    @property
    def n(self) -> int:
        return self._call("n")

    @property
    def n_residues(self) -> int:
        return self._call("n_residues")

    @property
    def n_chains(self) -> int:
        return self._call("n_chains")

    @property
    def model(self) -> int:
        return self._call("model")

    @property
    def chain_labels(self) -> bytes:
        return self._call("chain_labels")

    @property
    def classifier_name(self) -> bytes:
        return self._call("classifier_name")

    def xyz(self) -> np.ndarray:
        n = self.n
        return np.ctypeslib.as_array(self._call("coord_array"), shape=(n, 3)).copy() if n else np.zeros((0, 3))

    def radii(self) -> np.ndarray:
        n = self.n
        return np.ctypeslib.as_array(self._call("radius"), shape=(n,)).copy() if n else np.zeros(0)

    def add_atom(self, name: bytes, res_name: bytes, res_number: bytes, chain: bytes, x, y, z, classifier=None, options=None):
        if options is None and classifier is None:
            return self._call("add_atom", name, res_name, res_number, chain, x, y, z)
        return self._call("add_atom_wopt", name, res_name, res_number, chain, x, y, z, classifier, options or 0)

    def get_chains(self, chains: bytes, classifier=None, options: int = 0):
        h = self._call("get_chains", chains, classifier, options)
        return Structure(self.api, h) if h else None

    def atoms(self):
        """Everything the accessors expose per atom, as a list of tuples (for equality tests)."""
        c = self._call
        return [(c("atom_name", i), c("atom_res_name", i), c("atom_res_number", i), c("atom_chain_lcl", i), c("atom_symbol", i),
                 c("atom_class", i), c("atom_radius", i), c("atom_pdb_line", i)) for i in range(self.n)]

    def residues(self):
        out, first, last = [], ctypes.c_int(), ctypes.c_int()
        for r in range(self.n_residues):
            self._call("residue_atoms", r, ctypes.byref(first), ctypes.byref(last))
            ref = self._call("residue_reference", r)
            out.append((self._call("residue_name", r), self._call("residue_number", r), first.value, last.value,
                        (ref.contents.name, ref.contents.values()) if ref else None))
        return out

    def chains(self):
        out, a0, a1, r0, r1 = [], ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        for k in range(self.n_chains):
            label = self._call("chain_label", k)
            self._call("chain_atoms", label[:1] or b"\0", ctypes.byref(a0), ctypes.byref(a1))
            self._call("chain_residues", label[:1] or b"\0", ctypes.byref(r0), ctypes.byref(r1))
            out.append((label, a0.value, a1.value, r0.value, r1.value))
        return out

    def calc(self, parameters=None):
        """freesasa_calc_structure(); returns (per-atom SASA, total) or raises where C returns NULL."""
        L = self.api.lib
        res = L.freesasa_calc_structure(self.h, ctypes.byref(parameters) if parameters is not None else None)
        if not res:
            raise RuntimeError("freesasa_calc_structure returned NULL")
        sasa = np.ctypeslib.as_array(res.contents.sasa, shape=(self.n,)).copy()
        total = float(res.contents.total)
        L.freesasa_result_free(res)
        return sasa, total


_api = None


def api() -> StructureAPI:
    """The binding on this repo's libfreesasa_b200_host.so."""
    global _api
    if _api is None:
        from . import _CResult, Parameters, _host_lib

        _api = StructureAPI(_host_lib(), _CResult, Parameters)
    return _api


# ------------------------------------------------------------------------------------------------------------
# result tree (scope row f-3): ctypes view of the node API, reference src/freesasa.h:1449-1850
# ------------------------------------------------------------------------------------------------------------
NODE_ATOM, NODE_RESIDUE, NODE_CHAIN, NODE_STRUCTURE, NODE_RESULT, NODE_ROOT = range(6)


class TreeAPI:
    """The reference's tree entry points bound on ``api.lib``; ``walk`` flattens a tree for comparisons."""

    def __init__(self, api: StructureAPI):
        self.api, L = api, api.lib
        cp, ci, cd = ctypes.c_char_p, ctypes.c_int, ctypes.c_double
        res_p, par_p = ctypes.POINTER(api.Result), ctypes.POINTER(api.Parameters)

        def sig(name, restype, *argtypes):
            f = getattr(L, name)
            f.restype, f.argtypes = restype, list(argtypes)

        sig("freesasa_tree_new", _vp)
        sig("freesasa_tree_init", _vp, res_p, _vp, cp)
        sig("freesasa_tree_add_result", ci, _vp, res_p, _vp, cp)
        sig("freesasa_tree_join", ci, _vp, ctypes.POINTER(_vp))
        sig("freesasa_calc_tree", _vp, _vp, par_p, cp)
        sig("freesasa_node_free", ci, _vp)
        sig("freesasa_result_classes", NodeArea, _vp, res_p)
        sig("freesasa_node_area", ctypes.POINTER(NodeArea), _vp)
        for name in ("children", "next", "parent"):
            sig("freesasa_node_" + name, _vp, _vp)
        sig("freesasa_node_type", ci, _vp)
        for name in ("name", "classified_by", "atom_pdb_line", "atom_residue_number", "atom_residue_name", "atom_chain",
                     "residue_number", "structure_chain_labels"):
            sig("freesasa_node_" + name, cp, _vp)
        for name in ("atom_is_polar", "atom_is_mainchain", "residue_n_atoms", "chain_n_residues", "structure_n_chains",
                     "structure_n_atoms", "structure_model"):
            sig("freesasa_node_" + name, ci, _vp)
        sig("freesasa_node_atom_radius", cd, _vp)
        sig("freesasa_node_residue_reference", ctypes.POINTER(NodeArea), _vp)
        sig("freesasa_node_structure_result", res_p, _vp)
        sig("freesasa_node_result_parameters", par_p, _vp)

    def make_result(self, sasa: np.ndarray, parameters=None):
        """A caller-owned freesasa_result over ``sasa`` (kept alive by the returned tuple)."""
        sasa = np.ascontiguousarray(sasa, dtype=np.float64)
        r = self.api.Result()
        r.total = float(np.cumsum(sasa)[-1]) if sasa.size else 0.0  # serial sum in atom order, src/freesasa.c:113-116
        r.sasa = sasa.ctypes.data_as(_dp)
        r.n_atoms = int(sasa.shape[0])
        r.parameters = parameters if parameters is not None else self.api.Parameters(0, 1.4, 100, 20, 1)
        return r, sasa

    def init(self, result, structure: Structure, name: bytes):
        return self.api.lib.freesasa_tree_init(ctypes.byref(result), structure.h, name)

    def free(self, root):
        return self.api.lib.freesasa_node_free(root)

    def classes(self, structure: Structure, result):
        a = self.api.lib.freesasa_result_classes(structure.h, ctypes.byref(result))
        return (a.name,) + tuple(np.float64(v).view(np.uint64).item() for v in a.values())

    def walk(self, node, depth=0, out=None):
        """Pre-order list of (depth, type, name, area bits, properties) for the subtree under ``node``."""
        L = self.api.lib
        out = [] if out is None else out
        bits = lambda v: np.float64(v).view(np.uint64).item()  # noqa: E731
        t = L.freesasa_node_type(node)
        area = None
        if t not in (NODE_ROOT, NODE_RESULT):
            a = L.freesasa_node_area(node).contents
            area = (a.name,) + tuple(bits(v) for v in a.values())
        if t == NODE_ATOM:
            props = (L.freesasa_node_atom_is_polar(node), L.freesasa_node_atom_is_mainchain(node),
                     bits(L.freesasa_node_atom_radius(node)), L.freesasa_node_atom_pdb_line(node),
                     L.freesasa_node_atom_residue_number(node), L.freesasa_node_atom_residue_name(node),
                     L.freesasa_node_atom_chain(node))
        elif t == NODE_RESIDUE:
            ref = L.freesasa_node_residue_reference(node)
            props = (L.freesasa_node_residue_n_atoms(node), L.freesasa_node_residue_number(node),
                     (ref.contents.name, ref.contents.values()) if ref else None)
        elif t == NODE_CHAIN:
            props = (L.freesasa_node_chain_n_residues(node),)
        elif t == NODE_STRUCTURE:
            res = L.freesasa_node_structure_result(node).contents
            n = res.n_atoms
            props = (L.freesasa_node_structure_n_chains(node), L.freesasa_node_structure_n_atoms(node),
                     L.freesasa_node_structure_model(node), L.freesasa_node_structure_chain_labels(node), bits(res.total), n,
                     np.ctypeslib.as_array(res.sasa, shape=(n,)).view(np.uint64).tolist())
        elif t == NODE_RESULT:
            p = L.freesasa_node_result_parameters(node).contents
            props = (L.freesasa_node_classified_by(node), p.alg, p.probe_radius, p.shrake_rupley_n_points,
                     p.lee_richards_n_slices, p.n_threads)
        else:
            props = ()
        parent = L.freesasa_node_parent(node)
        out.append((depth, t, L.freesasa_node_name(node), area, props, bool(parent)))
        child = L.freesasa_node_children(node)
        while child:
            self.walk(child, depth + 1, out)
            child = L.freesasa_node_next(child)
        return out

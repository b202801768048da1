"""Seeded synthetic inputs for the benchmark configurations in BASELINE.json / SURVEY.md §8(d).

numpy only — usable without a GPU and without the CUDA library.  Every generator returns
``xyz`` as float64 ``[n, 3]`` (the reference's AoS coordinate layout, src/coord.h:26-38) and
``radii`` as float64 ``[n]`` (van der Waals radii WITHOUT the probe, as freesasa_calc_coord takes
them, src/freesasa.c:122-142).
"""
from __future__ import annotations

import numpy as np

# ProtOr-like heavy-atom radii (C aliphatic x3, C aromatic/carbonyl, N, N+, O, O-, S) — the ten
# entry table from SURVEY.md §8(d).
RADIUS_TABLE = np.array([1.88, 1.88, 1.88, 1.61, 1.76, 1.64, 1.64, 1.42, 1.46, 1.77])
LATTICE = 2.6  # Å; 17.6 Å^3 per atom ~ protein heavy-atom density
JITTER = 0.6  # Å, uniform +-


def _lattice_in_ball(r_out: float, r_in: float = 0.0):
    m = int(np.ceil(r_out / LATTICE)) + 1
    ax = np.arange(-m, m + 1, dtype=np.float64) * LATTICE
    g = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), axis=-1).reshape(-1, 3)  # scan order i,j,k
    d = np.sqrt((g * g).sum(1))
    keep = (d <= r_out) & (d >= r_in)
    return g[keep], d[keep]


def globule(n: int, seed: int = 0, offset=(0.0, 0.0, 0.0), shuffle: bool = False):
    """Dense globular pseudo-protein of exactly ``n`` atoms: jittered cubic lattice clipped to a
    ball (configs C2/C3 with n = 100000; the unit of C4 with n ~ 5000)."""
    rng = np.random.default_rng(np.random.PCG64(0x9E3779B97F4A7C15 ^ (seed * 0x100000001B3)))
    r = (3.0 * n * LATTICE**3 / (4.0 * np.pi)) ** (1.0 / 3.0)
    pts, d = _lattice_in_ball(r + 2 * LATTICE)
    order = np.argsort(d, kind="stable")[:n]
    order.sort()  # back to lattice scan order
    xyz = pts[order] + rng.uniform(-JITTER, JITTER, size=(n, 3))
    radii = RADIUS_TABLE[rng.integers(0, len(RADIUS_TABLE), size=n)]
    if shuffle:
        p = rng.permutation(n)
        xyz, radii = xyz[p], radii[p]
    xyz = xyz + np.asarray(offset, dtype=np.float64)
    return np.ascontiguousarray(xyz), np.ascontiguousarray(radii)


def capsid(n: int = 1_000_000, r_out: float = 250.0, seed: int = 0):
    """Hollow shell of ``n`` atoms (config C5): the same jittered lattice restricted to
    r_in <= r <= r_out with the shell volume = n * 17.58 Å^3."""
    rng = np.random.default_rng(np.random.PCG64(0xC2B2AE3D27D4EB4F ^ (seed * 0x100000001B3)))
    vol = n * LATTICE**3
    r_in3 = r_out**3 - 3.0 * vol / (4.0 * np.pi)
    r_in = max(r_in3, 0.0) ** (1.0 / 3.0)
    pts, d = _lattice_in_ball(r_out + LATTICE, max(r_in - 2 * LATTICE, 0.0))
    order = np.argsort(-d, kind="stable")[:n]  # outermost n lattice points
    order.sort()
    xyz = pts[order] + rng.uniform(-JITTER, JITTER, size=(len(order), 3))
    radii = RADIUS_TABLE[rng.integers(0, len(RADIUS_TABLE), size=len(order))]
    return np.ascontiguousarray(xyz), np.ascontiguousarray(radii)


def batch(n_struct: int, n_lo: int = 4000, n_hi: int = 6000, seed: int = 0):
    """``n_struct`` independent globules with sizes ~U[n_lo, n_hi] and a random rigid offset each
    (config C4).  Returns a list of (xyz, radii)."""
    rng = np.random.default_rng(np.random.PCG64(0x165667B19E3779F9 ^ (seed * 0x100000001B3)))
    out = []
    for k in range(n_struct):
        nk = int(rng.integers(n_lo, n_hi + 1))
        off = rng.uniform(-300.0, 300.0, size=3)
        out.append(globule(nk, seed=seed * 100003 + k + 1, offset=off))
    return out


# ---------------------------------------------------------------------------------------------------------
# PDB text (scope row f-1: ingest).  Synthetic but format-faithful ATOM/HETATM records: heavy atoms of the
# twenty standard residues in wwPDB naming and column layout, placed on the globule's coordinates (rounded to
# the format's three decimals), so that the same text exercises the reader AND gives a realistic SASA problem.
# ---------------------------------------------------------------------------------------------------------
RESIDUE_ATOMS = {
    "ALA": "N CA C O CB", "ARG": "N CA C O CB CG CD NE CZ NH1 NH2", "ASN": "N CA C O CB CG OD1 ND2",
    "ASP": "N CA C O CB CG OD1 OD2", "CYS": "N CA C O CB SG", "GLN": "N CA C O CB CG CD OE1 NE2",
    "GLU": "N CA C O CB CG CD OE1 OE2", "GLY": "N CA C O", "HIS": "N CA C O CB CG ND1 CD2 CE1 NE2",
    "ILE": "N CA C O CB CG1 CG2 CD1", "LEU": "N CA C O CB CG CD1 CD2", "LYS": "N CA C O CB CG CD CE NZ",
    "MET": "N CA C O CB CG SD CE", "PHE": "N CA C O CB CG CD1 CD2 CE1 CE2 CZ", "PRO": "N CA C O CB CG CD",
    "SER": "N CA C O CB OG", "THR": "N CA C O CB OG1 CG2", "TRP": "N CA C O CB CG CD1 CD2 NE1 CE2 CE3 CZ2 CZ3 CH2",
    "TYR": "N CA C O CB CG CD1 CD2 CE1 CE2 CZ OH", "VAL": "N CA C O CB CG1 CG2",
}
_RES_NAMES = sorted(RESIDUE_ATOMS)


def pdb_atom_line(serial, name, res_name, chain, res_seq, x, y, z, element, record="ATOM", alt=" ", icode=" ",
                  occupancy=1.0, bfactor=0.0):
    """One 80-column coordinate record (wwPDB format v3.3, ATOM/HETATM)."""
    nm = name if len(name) == 4 or len(element) == 2 else " " + name
    return ("%-6s%5d %-4s%1s%3s %1s%4d%1s   %8.3f%8.3f%8.3f%6.2f%6.2f          %2s  " %
            (record, serial % 100000, nm, alt, res_name, chain, res_seq % 10000, icode, x, y, z, occupancy, bfactor,
             element.rjust(2)))


def pdb_text(n_atoms: int, seed: int = 0, chains: int = 1, models: int = 1, hydrogens: float = 0.0, hetatm: int = 0,
             altloc: float = 0.0, unknown: float = 0.0, offset=(0.0, 0.0, 0.0), newline: str = "\n",
             element_column: bool = True, shuffle: bool = False) -> str:
    """PDB text with ~``n_atoms`` heavy protein atoms per model.

    chains      number of chains (labels A, B, ...), residues dealt to them in contiguous blocks
    models      MODEL/ENDMDL blocks (coordinates jittered per model), 1 = no MODEL records
    hydrogens   probability of a hydrogen record after a heavy atom (the reader drops them by default)
    hetatm      number of HETATM water oxygens appended per model
    altloc      probability that an atom is given as two alternate locations A/B (only A must survive)
    unknown     probability that a residue is a non-standard one ("UNK"/"LIG" with odd atom names)
    element_column  False: records are truncated after the B factor (no element symbol, 66 columns)
    shuffle     atoms take globule positions in random instead of lattice scan order
    """
    rng = np.random.default_rng(np.random.PCG64(0x27D4EB2F165667C5 ^ (seed * 0x100000001B3)))
    xyz, _ = globule(max(n_atoms, 8), seed=seed, offset=offset, shuffle=shuffle)
    out = ["HEADER    SYNTHETIC GLOBULE                        01-JAN-00   XXXX",
           "REMARK   1 generated by freesasa_b200.workloads.pdb_text"]
    labels = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789"
    for m in range(models):
        if models > 1:
            out.append("MODEL     %4d" % (m + 1))
        jitter = rng.normal(0, 0.05, size=xyz.shape) if m else 0.0
        pos = xyz + jitter
        serial, k, res_seq = 1, 0, 0
        per_chain = -(-len(pos) // chains)
        while k < len(pos):
            chain = labels[min(k // per_chain, chains - 1) % len(labels)]
            res_seq += 1
            if unknown and rng.random() < unknown:
                res_name, atoms = ("UNK", ["N", "CA", "C", "O", "CB", "XX1"]) if rng.random() < 0.5 else ("LIG", ["C1", "O1", "FE", "CL1", "N1"])
            else:
                res_name = _RES_NAMES[int(rng.integers(0, len(_RES_NAMES)))]
                atoms = RESIDUE_ATOMS[res_name].split()
            for name in atoms:
                if k >= len(pos) or (k and k % per_chain == 0 and name != atoms[0]):
                    break
                element = "FE" if name == "FE" else "CL" if name.startswith("CL") else name[0]
                x, y, z = pos[k]
                two = altloc and rng.random() < altloc
                rec = pdb_atom_line(serial, name, res_name, chain, res_seq, x, y, z, element, alt="A" if two else " ",
                                    occupancy=0.5 if two else 1.0, bfactor=float(rng.uniform(5, 60)))
                out.append(rec if element_column else rec[:66])
                serial += 1
                if two:
                    rec = pdb_atom_line(serial, name, res_name, chain, res_seq, x + 0.3, y - 0.2, z + 0.1, element, alt="B",
                                        occupancy=0.5, bfactor=20.0)
                    out.append(rec if element_column else rec[:66])
                    serial += 1
                if hydrogens and rng.random() < hydrogens:
                    hname = ("H" + name[1:])[:4] if len(name) > 1 else "H"
                    out.append(pdb_atom_line(serial, hname, res_name, chain, res_seq, x + 0.6, y + 0.6, z + 0.6, "H"))
                    serial += 1
                k += 1
            if k < len(pos) and k % per_chain == 0:
                out.append("TER   %5d      %3s %1s%4d" % (serial % 100000, res_name, chain, res_seq % 10000))
                serial += 1
        for w in range(hetatm):
            p = rng.uniform(-1, 1, size=3) * (pos.max() - pos.min()) / 2 + pos.mean(0)
            out.append(pdb_atom_line(serial, "O", "HOH", chain, res_seq + 1 + w, p[0], p[1], p[2], "O", record="HETATM"))
            serial += 1
        if models > 1:
            out.append("ENDMDL")
    out.append("END")
    return newline.join(out) + newline

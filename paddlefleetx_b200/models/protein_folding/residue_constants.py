"""Amino-acid chemistry tables used by the structure features (reference ppfleetx/models/protein_folding/residue_constants.py:1-961).

Everything here is public protein chemistry (the 20 standard residues, the 37-name heavy-atom vocabulary of PDB files, side-chain
torsion definitions, the dense 14-slot per-residue atom layout).  Index maps between the 37- and 14-slot layouts are derived at
import time instead of being listed.  The reference's ideal-geometry table for rebuilding side chains from torsion angles
(``rigid_group_atom_positions``) belongs to the structure module, which lives outside this repository's scope (as it does in the
reference: the folding head is in PaddleHelix), and is not included.
"""
from __future__ import annotations

import os
from typing import Dict, List, NamedTuple

import numpy as np

# one-letter codes in the canonical order (alphabetical by three-letter code); index 20 = unknown 'X'
restypes: List[str] = ["A", "R", "N", "D", "C", "Q", "E", "G", "H", "I", "L", "K", "M", "F", "P", "S", "T", "W", "Y", "V"]
restype_order: Dict[str, int] = {r: i for i, r in enumerate(restypes)}
restype_num = len(restypes)
unk_restype_index = restype_num
restypes_with_x = restypes + ["X"]
restype_order_with_x = {r: i for i, r in enumerate(restypes_with_x)}
restypes_with_x_and_gap = restypes_with_x + ["-"]

restype_1to3 = {"A": "ALA", "R": "ARG", "N": "ASN", "D": "ASP", "C": "CYS", "Q": "GLN", "E": "GLU", "G": "GLY", "H": "HIS", "I": "ILE", "L": "LEU",
                "K": "LYS", "M": "MET", "F": "PHE", "P": "PRO", "S": "SER", "T": "THR", "W": "TRP", "Y": "TYR", "V": "VAL"}
restype_3to1 = {v: k for k, v in restype_1to3.items()}
unk_restype = "UNK"
resnames = [restype_1to3[r] for r in restypes] + [unk_restype]
resname_to_idx = {n: i for i, n in enumerate(resnames)}

# the 37 heavy-atom names that occur in the standard residues (+ terminal OXT); positions in this list are the "atom37" slots
atom_types: List[str] = ["N", "CA", "C", "CB", "O", "CG", "CG1", "CG2", "OG", "OG1", "SG", "CD", "CD1", "CD2", "ND1", "ND2", "OD1", "OD2", "SD",
                         "CE", "CE1", "CE2", "CE3", "NE", "NE1", "NE2", "OE1", "OE2", "CH2", "NH1", "NH2", "OH", "CZ", "CZ2", "CZ3", "NZ", "OXT"]
atom_order: Dict[str, int] = {a: i for i, a in enumerate(atom_types)}
atom_type_num = len(atom_types)

# dense per-residue layout: at most 14 heavy atoms, backbone first
restype_name_to_atom14_names: Dict[str, List[str]] = {
    "ALA": ["N", "CA", "C", "O", "CB"],
    "ARG": ["N", "CA", "C", "O", "CB", "CG", "CD", "NE", "CZ", "NH1", "NH2"],
    "ASN": ["N", "CA", "C", "O", "CB", "CG", "OD1", "ND2"],
    "ASP": ["N", "CA", "C", "O", "CB", "CG", "OD1", "OD2"],
    "CYS": ["N", "CA", "C", "O", "CB", "SG"],
    "GLN": ["N", "CA", "C", "O", "CB", "CG", "CD", "OE1", "NE2"],
    "GLU": ["N", "CA", "C", "O", "CB", "CG", "CD", "OE1", "OE2"],
    "GLY": ["N", "CA", "C", "O"],
    "HIS": ["N", "CA", "C", "O", "CB", "CG", "ND1", "CD2", "CE1", "NE2"],
    "ILE": ["N", "CA", "C", "O", "CB", "CG1", "CG2", "CD1"],
    "LEU": ["N", "CA", "C", "O", "CB", "CG", "CD1", "CD2"],
    "LYS": ["N", "CA", "C", "O", "CB", "CG", "CD", "CE", "NZ"],
    "MET": ["N", "CA", "C", "O", "CB", "CG", "SD", "CE"],
    "PHE": ["N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "CE1", "CE2", "CZ"],
    "PRO": ["N", "CA", "C", "O", "CB", "CG", "CD"],
    "SER": ["N", "CA", "C", "O", "CB", "OG"],
    "THR": ["N", "CA", "C", "O", "CB", "OG1", "CG2"],
    "TRP": ["N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "NE1", "CE2", "CE3", "CZ2", "CZ3", "CH2"],
    "TYR": ["N", "CA", "C", "O", "CB", "CG", "CD1", "CD2", "CE1", "CE2", "CZ", "OH"],
    "VAL": ["N", "CA", "C", "O", "CB", "CG1", "CG2"],
    "UNK": [],
}
for _names in restype_name_to_atom14_names.values():
    _names.extend([""] * (14 - len(_names)))
residue_atoms: Dict[str, List[str]] = {k: sorted(a for a in v if a) for k, v in restype_name_to_atom14_names.items() if k != "UNK"}

# side-chain torsions chi1..chi4: the four atoms that define each
chi_angles_atoms: Dict[str, List[List[str]]] = {
    "ALA": [],
    "ARG": [["N", "CA", "CB", "CG"], ["CA", "CB", "CG", "CD"], ["CB", "CG", "CD", "NE"], ["CG", "CD", "NE", "CZ"]],
    "ASN": [["N", "CA", "CB", "CG"], ["CA", "CB", "CG", "OD1"]],
    "ASP": [["N", "CA", "CB", "CG"], ["CA", "CB", "CG", "OD1"]],
    "CYS": [["N", "CA", "CB", "SG"]],
    "GLN": [["N", "CA", "CB", "CG"], ["CA", "CB", "CG", "CD"], ["CB", "CG", "CD", "OE1"]],
    "GLU": [["N", "CA", "CB", "CG"], ["CA", "CB", "CG", "CD"], ["CB", "CG", "CD", "OE1"]],
    "GLY": [],
    "HIS": [["N", "CA", "CB", "CG"], ["CA", "CB", "CG", "ND1"]],
    "ILE": [["N", "CA", "CB", "CG1"], ["CA", "CB", "CG1", "CD1"]],
    "LEU": [["N", "CA", "CB", "CG"], ["CA", "CB", "CG", "CD1"]],
    "LYS": [["N", "CA", "CB", "CG"], ["CA", "CB", "CG", "CD"], ["CB", "CG", "CD", "CE"], ["CG", "CD", "CE", "NZ"]],
    "MET": [["N", "CA", "CB", "CG"], ["CA", "CB", "CG", "SD"], ["CB", "CG", "SD", "CE"]],
    "PHE": [["N", "CA", "CB", "CG"], ["CA", "CB", "CG", "CD1"]],
    "PRO": [["N", "CA", "CB", "CG"], ["CA", "CB", "CG", "CD"]],
    "SER": [["N", "CA", "CB", "OG"]],
    "THR": [["N", "CA", "CB", "OG1"]],
    "TRP": [["N", "CA", "CB", "CG"], ["CA", "CB", "CG", "CD1"]],
    "TYR": [["N", "CA", "CB", "CG"], ["CA", "CB", "CG", "CD1"]],
    "VAL": [["N", "CA", "CB", "CG1"]],
}
# [21, 4] per residue type (+ UNK row): which chi angles exist / which are symmetric under a 180-degree flip of the terminal group
chi_angles_mask: List[List[float]] = [[1.0 if k < len(chi_angles_atoms[restype_1to3[r]]) else 0.0 for k in range(4)] for r in restypes] + [[0.0] * 4]
_PI_PERIODIC = {"ASP": 1, "GLU": 2, "PHE": 1, "TYR": 1}      # chi index (0-based) whose terminal atoms are chemically equivalent
chi_pi_periodic: List[List[float]] = [[1.0 if _PI_PERIODIC.get(restype_1to3[r], -1) == k else 0.0 for k in range(4)] for r in restypes] + [[0.0] * 4]

# atom pairs whose names can be swapped without changing the molecule (naming ambiguity of symmetric side chains)
residue_atom_renaming_swaps: Dict[str, Dict[str, str]] = {
    "ASP": {"OD1": "OD2"}, "GLU": {"OE1": "OE2"}, "PHE": {"CD1": "CD2", "CE1": "CE2"}, "TYR": {"CD1": "CD2", "CE1": "CE2"}}

van_der_waals_radius = {"C": 1.7, "N": 1.55, "O": 1.52, "S": 1.8}
ca_ca = 3.80209737096                                   # ideal CA-CA distance of consecutive trans residues (Angstrom)
between_res_bond_length_c_n = [1.329, 1.341]            # peptide bond C-N: general, proline
between_res_bond_length_stddev_c_n = [0.014, 0.016]
between_res_cos_angles_c_n_ca = [-0.5203, 0.0353]       # cos(C-N-CA): mean, stddev
between_res_cos_angles_ca_c_n = [-0.4473, 0.0311]       # cos(CA-C-N): mean, stddev


def _make_atom14_maps():
    a14_to_37 = np.zeros((restype_num + 1, 14), np.int64)
    a37_to_14 = np.zeros((restype_num + 1, atom_type_num), np.int64)
    m14 = np.zeros((restype_num + 1, 14), np.float32)
    m37 = np.zeros((restype_num + 1, atom_type_num), np.float32)
    for i, name in enumerate(resnames):
        for slot, atom in enumerate(restype_name_to_atom14_names[name]):
            if atom:
                a14_to_37[i, slot], a37_to_14[i, atom_order[atom]] = atom_order[atom], slot
                m14[i, slot] = m37[i, atom_order[atom]] = 1.0
    return a14_to_37, a37_to_14, m14, m37


restype_atom14_to_atom37, restype_atom37_to_atom14, restype_atom14_mask, restype_atom37_mask = _make_atom14_maps()


def sequence_to_onehot(sequence: str, mapping: Dict[str, int] = None, map_unknown_to_x: bool = True) -> np.ndarray:
    mapping = restype_order_with_x if mapping is None else mapping
    n = max(mapping.values()) + 1
    out = np.zeros((len(sequence), n), np.int32)
    for i, aa in enumerate(sequence):
        if aa not in mapping:
            if not (map_unknown_to_x and "X" in mapping):
                raise ValueError(f"invalid residue {aa!r} at position {i}")
            aa = "X"
        out[i, mapping[aa]] = 1
    return out


def aatype_to_str_sequence(aatype) -> str:
    return "".join(restypes_with_x[int(i)] for i in aatype)


# ---------------------------------------------------------------------------------------------------------------- stereo-chemistry tables
class Bond(NamedTuple):
    atom1_name: str
    atom2_name: str
    length: float
    stddev: float


class BondAngle(NamedTuple):
    atom1_name: str
    atom2_name: str
    atom3name: str
    angle_rad: float
    stddev: float


STEREO_CHEMICAL_PROPS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stereo_chemical_props.txt")


def load_stereo_chemical_props(path: str = None):
    """Literature bond lengths and angles per residue type, read from ``stereo_chemical_props.txt`` (the Engh & Huber table that ships with
    AlphaFold's parameters: a ``Bond Residue Mean StdDev`` block, a line ``-``, an ``Angle Residue Mean StdDev`` block in degrees, a line
    ``-``).  Returns ``(residue_bonds, residue_virtual_bonds, residue_bond_angles)``; a "virtual bond" is the 1-3 distance an angle implies by
    the law of cosines, with the error propagated from the two bond lengths and the angle (reference residue_constants.py:403-488).  The table
    is data, not code, and is not in the reference tree either: put it next to this module or pass ``path``."""
    path = path or STEREO_CHEMICAL_PROPS
    if not os.path.isfile(path):
        raise FileNotFoundError(f"{path} not found: copy stereo_chemical_props.txt (AlphaFold parameter release) there, or pass path=")
    with open(path, "rt") as f:
        lines = iter(f.read().splitlines())
    residue_bonds: Dict[str, List[Bond]] = {}
    next(lines)                                            # header
    for line in lines:
        if line.strip() == "-":
            break
        bond, resname, length, stddev = line.split()
        a1, a2 = bond.split("-")
        residue_bonds.setdefault(resname, []).append(Bond(a1, a2, float(length), float(stddev)))
    residue_bonds["UNK"] = []
    residue_bond_angles: Dict[str, List[BondAngle]] = {}
    next(lines)                                            # blank
    next(lines)                                            # header
    for line in lines:
        if line.strip() == "-":
            break
        bond, resname, angle, stddev = line.split()
        a1, a2, a3 = bond.split("-")
        residue_bond_angles.setdefault(resname, []).append(BondAngle(a1, a2, a3, float(angle) / 180.0 * np.pi, float(stddev) / 180.0 * np.pi))
    residue_bond_angles["UNK"] = []
    residue_virtual_bonds: Dict[str, List[Bond]] = {}
    for resname, angles in residue_bond_angles.items():
        by_pair = {"-".join(sorted((b.atom1_name, b.atom2_name))): b for b in residue_bonds.get(resname, [])}
        virtual = []
        for ba in angles:
            b1 = by_pair["-".join(sorted((ba.atom1_name, ba.atom2_name)))]
            b2 = by_pair["-".join(sorted((ba.atom2_name, ba.atom3name)))]
            g = ba.angle_rad
            length = np.sqrt(b1.length ** 2 + b2.length ** 2 - 2 * b1.length * b2.length * np.cos(g))
            outer = 0.5 / length                                           # d sqrt(u) / du
            d_gamma = 2 * b1.length * b2.length * np.sin(g) * outer
            d_b1 = (2 * b1.length - 2 * b2.length * np.cos(g)) * outer
            d_b2 = (2 * b2.length - 2 * b1.length * np.cos(g)) * outer
            stddev = np.sqrt((d_gamma * ba.stddev) ** 2 + (d_b1 * b1.stddev) ** 2 + (d_b2 * b2.stddev) ** 2)
            virtual.append(Bond(ba.atom1_name, ba.atom3name, float(length), float(stddev)))
        residue_virtual_bonds[resname] = virtual
    return residue_bonds, residue_virtual_bonds, residue_bond_angles


def chi_angle_atom(atom_index: int) -> np.ndarray:
    """``[21, 37, 4]`` one-hot: for every residue type and chi angle, which atom37 slot holds the ``atom_index``-th of the four atoms defining
    that torsion (all-zero columns where the residue has fewer chi angles; reference residue_constants.py:757-776)."""
    out = np.zeros((restype_num + 1, atom_type_num, 4), np.float64)
    for r, letter in enumerate(restypes):
        for k, atoms in enumerate(chi_angles_atoms[restype_1to3[letter]]):
            out[r, atom_order[atoms[atom_index]], k] = 1.0
    return out


chi_atom_1_one_hot = chi_angle_atom(1)
chi_atom_2_one_hot = chi_angle_atom(2)


def make_atom14_dists_bounds(overlap_tolerance: float = 1.5, bond_length_tolerance_factor: float = 15, props=None):
    """Per residue type, ``[21, 14, 14]`` lower / upper bounds on intra-residue atom distances used to flag structural violations: non-bonded
    pairs may not come closer than the sum of their van-der-Waals radii minus ``overlap_tolerance``; bonded and 1-3 ("virtual bond") pairs must
    stay within ``bond_length_tolerance_factor`` standard deviations of the literature length (reference residue_constants.py:908-961).
    ``props``: the triple from ``load_stereo_chemical_props`` (loaded from the default file when omitted)."""
    lower = np.zeros((restype_num + 1, 14, 14), np.float32)
    upper = np.zeros((restype_num + 1, 14, 14), np.float32)
    stddev = np.zeros((restype_num + 1, 14, 14), np.float32)
    residue_bonds, residue_virtual_bonds, _ = props if props is not None else load_stereo_chemical_props()
    for r, letter in enumerate(restypes):
        resname = restype_1to3[letter]
        names = restype_name_to_atom14_names[resname]
        for i, a in enumerate(names):
            for j, b in enumerate(names):
                if a and b and i != j:
                    lower[r, i, j] = van_der_waals_radius[a[0]] + van_der_waals_radius[b[0]] - overlap_tolerance
                    upper[r, i, j] = 1e10
        for bond in list(residue_bonds.get(resname, [])) + list(residue_virtual_bonds.get(resname, [])):
            i, j = names.index(bond.atom1_name), names.index(bond.atom2_name)
            lo, hi = bond.length - bond_length_tolerance_factor * bond.stddev, bond.length + bond_length_tolerance_factor * bond.stddev
            lower[r, i, j] = lower[r, j, i] = lo
            upper[r, i, j] = upper[r, j, i] = hi
            stddev[r, i, j] = stddev[r, j, i] = bond.stddev
    return {"lower_bound": lower, "upper_bound": upper, "stddev": stddev}

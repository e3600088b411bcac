"""All-atom structure features: backbone / side-chain torsion angles from atom37 coordinates (reference
ppfleetx/models/protein_folding/all_atom.py:25-254; consumed by the template torsion-angle embedding, evoformer.py:893-935).

One gather builds the ``[..., 7, 4, 3]`` tensor of the four atoms of each of the 7 torsions (pre-omega, phi, psi, chi1-4); the angle
is read off as the position of the fourth atom in the frame spanned by the first three, so the whole feature is a few batched
tensor ops with no Python loop over residues or angles.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import r3
from . import residue_constants as rc
from .common import no_autocast


def get_chi_atom_indices() -> List[List[List[int]]]:
    """[21 residue types (+UNK)][4 chi][4 atoms] -> atom37 indices, zero-filled where a chi angle does not exist."""
    out = []
    for r in rc.restypes:
        chis = [[rc.atom_order[a] for a in atoms] for atoms in rc.chi_angles_atoms[rc.restype_1to3[r]]]
        out.append(chis + [[0, 0, 0, 0]] * (4 - len(chis)))
    out.append([[0, 0, 0, 0]] * 4)
    return out


@no_autocast
def atom37_to_torsion_angles(aatype: torch.Tensor, all_atom_pos: torch.Tensor, all_atom_mask: torch.Tensor,
                             placeholder_for_undefined: bool = False) -> Dict[str, torch.Tensor]:
    """aatype ``[B, T, N]`` int, all_atom_pos ``[B, T, N, 37, 3]``, all_atom_mask ``[B, T, N, 37]`` ->

    ``torsion_angles_sin_cos`` / ``alt_torsion_angles_sin_cos`` ``[B, T, N, 7, 2]`` (sin, cos; the alternative set flips the pi-periodic
    chi angles by 180 degrees) and ``torsion_angles_mask`` ``[B, T, N, 7]``.  Order: pre-omega, phi, psi, chi1..chi4."""
    aatype = aatype.clamp(max=20)
    pos, mask = all_atom_pos, all_atom_mask.to(all_atom_pos.dtype)
    # previous residue's atoms (zero-padded at the first residue)
    prev_pos = torch.cat([torch.zeros_like(pos[..., :1, :, :]), pos[..., :-1, :, :]], dim=-3)
    prev_mask = torch.cat([torch.zeros_like(mask[..., :1, :]), mask[..., :-1, :]], dim=-2)
    N, CA, C, O = 0, 1, 2, 4
    pre_omega = torch.cat([prev_pos[..., [CA, C], :], pos[..., [N, CA], :]], dim=-2)               # CA(i-1) C(i-1) N(i) CA(i)
    phi = torch.cat([prev_pos[..., [C], :], pos[..., [N, CA, C], :]], dim=-2)                      # C(i-1) N CA C
    psi = pos[..., [N, CA, C, O], :]                                                               # N CA C O  (mirrored below: O is opposite N(i+1))
    pre_omega_mask = prev_mask[..., CA] * prev_mask[..., C] * mask[..., N] * mask[..., CA]
    phi_mask = prev_mask[..., C] * mask[..., N] * mask[..., CA] * mask[..., C]
    psi_mask = mask[..., N] * mask[..., CA] * mask[..., C] * mask[..., O]
    # side chains: gather the 4x4 atom indices of this residue type
    chi_idx = torch.as_tensor(get_chi_atom_indices(), device=pos.device)[aatype]                  # [B, T, N, 4, 4]
    flat = chi_idx.flatten(-2)                                                                     # [B, T, N, 16]
    chi_atoms = torch.gather(pos, -2, flat.unsqueeze(-1).expand(flat.shape + (3,))).unflatten(-2, (4, 4))
    chi_exists = torch.as_tensor(rc.chi_angles_mask, device=pos.device, dtype=pos.dtype)[aatype]   # [B, T, N, 4]
    chi_atom_mask = torch.gather(mask, -1, flat).unflatten(-1, (4, 4)).prod(-1)
    chis_mask = chi_exists * chi_atom_mask
    atoms = torch.cat([pre_omega.unsqueeze(-3), phi.unsqueeze(-3), psi.unsqueeze(-3), chi_atoms], dim=-3)   # [B, T, N, 7, 4, 3]
    torsion_mask = torch.cat([pre_omega_mask.unsqueeze(-1), phi_mask.unsqueeze(-1), psi_mask.unsqueeze(-1), chis_mask], dim=-1)
    # frame on atoms (1, 2 | 0): the torsion is the polar angle of atom 3 around the x axis of that frame
    frames = r3.rigids_from_3_points(point_on_neg_x_axis=atoms[..., 1, :], origin=atoms[..., 2, :], point_on_xy_plane=atoms[..., 0, :])
    fourth = r3.rigids_mul_vecs(r3.invert_rigids(frames), atoms[..., 3, :])
    sin_cos = torch.stack([fourth[..., 2], fourth[..., 1]], dim=-1)
    sin_cos = sin_cos / torch.sqrt((sin_cos * sin_cos).sum(-1, keepdim=True) + 1e-8)
    sin_cos = sin_cos * torch.tensor([1.0, 1.0, -1.0, 1.0, 1.0, 1.0, 1.0], device=pos.device, dtype=pos.dtype)[:, None]
    periodic = torch.as_tensor(rc.chi_pi_periodic, device=pos.device, dtype=pos.dtype)[aatype]
    flip = torch.cat([torch.ones_like(periodic[..., :3]), 1.0 - 2.0 * periodic], dim=-1)
    alt = sin_cos * flip.unsqueeze(-1)
    if placeholder_for_undefined:
        placeholder = torch.stack([torch.ones_like(sin_cos[..., 0]), torch.zeros_like(sin_cos[..., 1])], dim=-1)
        m = torsion_mask.unsqueeze(-1)
        sin_cos, alt = sin_cos * m + placeholder * (1 - m), alt * m + placeholder * (1 - m)
    return {"torsion_angles_sin_cos": sin_cos, "alt_torsion_angles_sin_cos": alt, "torsion_angles_mask": torsion_mask}

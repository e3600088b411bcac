"""``OuterProductMean`` (reference ppfleetx/models/protein_folding/outer_product_mean.py:23-150): the MSA -> pair update."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...distributed.protein_folding import dap


class OuterProductMean(nn.Module):
    def __init__(self, c_m, c_z, c_hidden=32):
        super().__init__()
        self.ln = nn.LayerNorm(c_m)
        self.left, self.right = nn.Linear(c_m, c_hidden), nn.Linear(c_m, c_hidden)
        self.out = nn.Linear(c_hidden * c_hidden, c_z)
        nn.init.zeros_(self.out.weight); nn.init.zeros_(self.out.bias)

    def forward(self, msa, msa_mask):
        # msa [b, S, R(/n), c] sharded by residues: left uses the local residues, right needs all residues
        mask = msa_mask.unsqueeze(-1)
        m = self.ln(msa)
        left = self.left(m) * mask
        right = dap.gather_full(self.right(m) * mask, axis=2)
        mask_full = dap.gather_full(mask, axis=2)
        outer = torch.einsum("bsic,bsjd->bijcd", left, right)
        norm = torch.einsum("bsic,bsjd->bijcd", mask, mask_full)
        out = self.out(outer.flatten(-2))
        return out / (norm.flatten(-2) + 1e-3)

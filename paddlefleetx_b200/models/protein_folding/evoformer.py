"""Evoformer stack of AlphaFold2/HelixFold with DAP / BP parallelism (reference models/protein_folding/{evoformer.py,
attentions.py,common.py,...}; distributed hooks evoformer.py:195-501).

Blocks: MSA row attention with pair bias, MSA column attention, MSA transition, outer-product-mean, triangle
multiplication (outgoing / incoming), triangle attention (starting / ending node), pair transition.  Gated attention runs
through SDPA (the reference's ``fused_gate_attention``, attentions.py:126-142).

Parallel layout under DAP (dap group size n): MSA activation ``[b, S/n, R, c]`` sharded by sequences for row-wise ops and
``[b, S, R/n, c]`` by residues for column-wise ops (``row_to_col`` / ``col_to_row`` all-to-all in between); pair activation
sharded along its first residue axis.  Under BP (size 2) rank 0 computes the MSA branch and rank 1 the pair branch of a
block, then ``sync_evoformer_results`` exchanges the outputs.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...distributed.protein_folding import bp, dap
from ...distributed.protein_folding.scg import scg


class GatedAttention(nn.Module):
    """Multi-head attention with sigmoid output gating and an optional additive bias ``[b, h, q, k]`` (per leading group)."""

    def __init__(self, q_dim, kv_dim, num_head, out_dim, gating=True):
        super().__init__()
        self.h, self.d = num_head, q_dim // num_head
        self.q, self.k, self.v = nn.Linear(q_dim, q_dim, bias=False), nn.Linear(kv_dim, q_dim, bias=False), nn.Linear(kv_dim, q_dim, bias=False)
        self.o = nn.Linear(q_dim, out_dim)
        self.g = nn.Linear(q_dim, q_dim) if gating else None
        if self.g is not None:
            nn.init.zeros_(self.g.weight); nn.init.ones_(self.g.bias)
        nn.init.zeros_(self.o.weight); nn.init.zeros_(self.o.bias)

    def forward(self, q_data, m_data, bias=None, nonbatched_bias=None):
        # q_data: [b, g, q, c]; m_data: [b, g, k, c]; bias: [b, g, 1, 1, k]; nonbatched_bias: [b, 1, h, q, k]
        b, g, nq, _ = q_data.shape
        nk = m_data.shape[2]
        q = self.q(q_data).view(b, g, nq, self.h, self.d).transpose(2, 3)
        k = self.k(m_data).view(b, g, nk, self.h, self.d).transpose(2, 3)
        v = self.v(m_data).view(b, g, nk, self.h, self.d).transpose(2, 3)
        mask = None
        if bias is not None:
            mask = bias.to(q.dtype)
        if nonbatched_bias is not None:
            mask = nonbatched_bias.to(q.dtype) if mask is None else mask + nonbatched_bias.to(q.dtype)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)
        o = o.transpose(2, 3).reshape(b, g, nq, self.h * self.d)
        if self.g is not None:
            o = o * torch.sigmoid(self.g(q_data))
        return self.o(o)


class MSARowAttentionWithPairBias(nn.Module):
    def __init__(self, c_m, c_z, num_head=8):
        super().__init__()
        self.ln_m, self.ln_z = nn.LayerNorm(c_m), nn.LayerNorm(c_z)
        self.pair_bias = nn.Linear(c_z, num_head, bias=False)
        self.attn = GatedAttention(c_m, c_m, num_head, c_m)

    def forward(self, msa, msa_mask, pair):
        # msa [b, S(/n), R, c_m]; pair [b, R(/n), R, c_z] sharded on dim 1 under DAP
        m = self.ln_m(msa)
        z = self.pair_bias(self.ln_z(pair))                       # [b, R/n, R, h]
        z = dap.all_gather(z, axis=1)                             # full [b, R, R, h] on every rank
        nb = z.permute(0, 3, 1, 2).unsqueeze(1)                   # [b, 1, h, R, R]
        bias = (1e9 * (msa_mask - 1.0))[:, :, None, None, :]
        return self.attn(m, m, bias, nb)


class MSAColumnAttention(nn.Module):
    def __init__(self, c_m, num_head=8):
        super().__init__()
        self.ln = nn.LayerNorm(c_m)
        self.attn = GatedAttention(c_m, c_m, num_head, c_m)

    def forward(self, msa, msa_mask):
        # column-wise: attend over sequences for every residue -> transpose S and R
        m = self.ln(msa).transpose(1, 2)
        mask = msa_mask.transpose(1, 2)
        bias = (1e9 * (mask - 1.0))[:, :, None, None, :]
        return self.attn(m, m, bias).transpose(1, 2)


class Transition(nn.Module):
    def __init__(self, c, factor=4):
        super().__init__()
        self.ln = nn.LayerNorm(c)
        self.fc1, self.fc2 = nn.Linear(c, c * factor), nn.Linear(c * factor, c)
        nn.init.zeros_(self.fc2.weight); nn.init.zeros_(self.fc2.bias)

    def forward(self, x, mask=None):
        return self.fc2(F.relu(self.fc1(self.ln(x))))


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
        right = dap.all_gather(self.right(m) * mask, axis=2)
        mask_full = dap.all_gather(mask, axis=2)
        outer = torch.einsum("bsic,bsjd->bijcd", left, right)
        norm = torch.einsum("bsic,bsjd->bijcd", mask, mask_full)
        out = self.out(outer.flatten(-2))
        return out / (norm.flatten(-2) + 1e-3)


class TriangleMultiplication(nn.Module):
    def __init__(self, c_z, c_hidden=128, outgoing=True):
        super().__init__()
        self.outgoing = outgoing
        self.ln_in, self.ln_out = nn.LayerNorm(c_z), nn.LayerNorm(c_hidden)
        self.left, self.right = nn.Linear(c_z, c_hidden), nn.Linear(c_z, c_hidden)
        self.left_gate, self.right_gate = nn.Linear(c_z, c_hidden), nn.Linear(c_z, c_hidden)
        self.out, self.gate = nn.Linear(c_hidden, c_z), nn.Linear(c_z, c_z)
        for g in (self.left_gate, self.right_gate, self.gate):
            nn.init.zeros_(g.weight); nn.init.ones_(g.bias)
        nn.init.zeros_(self.out.weight); nn.init.zeros_(self.out.bias)

    def forward(self, pair, pair_mask):
        # pair [b, R/n, R, c] (rows sharded).  outgoing: out[i,j] = sum_k a[i,k] b[j,k] -> needs all rows of b.
        mask = pair_mask.unsqueeze(-1)
        z = self.ln_in(pair)
        a = self.left(z) * mask * torch.sigmoid(self.left_gate(z))
        b_ = self.right(z) * mask * torch.sigmoid(self.right_gate(z))
        if self.outgoing:
            b_full = dap.all_gather(b_, axis=1)
            x = torch.einsum("bikc,bjkc->bijc", a, b_full)
        else:
            # incoming: out[i,j] = sum_k a[k,i] b[k,j]: contraction runs over the sharded axis -> work on column shards
            a_c, b_c = dap.row_to_col(a), dap.row_to_col(b_)         # [b, R, R/n, c]
            b_full = dap.all_gather(b_c, axis=2)
            x = torch.einsum("bkic,bkjc->bijc", a_c, b_full)          # [b, R/n(i), R, c]
        return self.out(self.ln_out(x)) * torch.sigmoid(self.gate(z))


class TriangleAttention(nn.Module):
    def __init__(self, c_z, num_head=4, starting=True):
        super().__init__()
        self.starting = starting
        self.ln = nn.LayerNorm(c_z)
        self.bias = nn.Linear(c_z, num_head, bias=False)
        self.attn = GatedAttention(c_z, c_z, num_head, c_z)

    def forward(self, pair, pair_mask):
        # starting node: attention along rows of the row-sharded pair; ending node: same on the transposed tensor
        if not self.starting:
            pair, pair_mask = dap.row_to_col(pair).transpose(1, 2), dap.row_to_col(pair_mask.unsqueeze(-1)).squeeze(-1).transpose(1, 2)
        z = self.ln(pair)
        nb = dap.all_gather(self.bias(z), axis=1).permute(0, 3, 1, 2).unsqueeze(1)
        bias = (1e9 * (pair_mask - 1.0))[:, :, None, None, :]
        out = self.attn(z, z, bias, nb)
        if not self.starting:
            out = dap.col_to_row(out.transpose(1, 2))
        return out


class EvoformerIteration(nn.Module):
    def __init__(self, c_m=256, c_z=128, msa_heads=8, pair_heads=4, dropout_msa=0.15, dropout_pair=0.25, is_extra_msa=False, outer_first=False):
        super().__init__()
        self.msa_row = MSARowAttentionWithPairBias(c_m, c_z, msa_heads)
        self.msa_col = MSAColumnAttention(c_m, msa_heads)
        self.msa_transition = Transition(c_m)
        self.outer = OuterProductMean(c_m, c_z)
        self.tri_mul_out, self.tri_mul_in = TriangleMultiplication(c_z, outgoing=True), TriangleMultiplication(c_z, outgoing=False)
        self.tri_att_start, self.tri_att_end = TriangleAttention(c_z, pair_heads, True), TriangleAttention(c_z, pair_heads, False)
        self.pair_transition = Transition(c_z)
        self.dm, self.dz = dropout_msa, dropout_pair
        self.outer_first = outer_first

    def _row_dropout(self, x, p, dim):
        if not self.training or p == 0:
            return x
        shape = list(x.shape)
        shape[dim] = 1
        return x * torch.bernoulli(torch.full(shape, 1 - p, device=x.device, dtype=x.dtype)) / (1 - p)

    def _msa_branch(self, msa, pair, msa_mask):
        msa = msa + self._row_dropout(self.msa_row(msa, msa_mask, pair), self.dm, 1)
        msa_c, mask_c = dap.row_to_col(msa), dap.row_to_col(msa_mask.unsqueeze(-1)).squeeze(-1)
        msa_c = msa_c + self.msa_col(msa_c, mask_c)
        msa_c = msa_c + self.msa_transition(msa_c)
        return msa_c, mask_c

    def _pair_branch(self, pair, pair_mask):
        pair = pair + self._row_dropout(self.tri_mul_out(pair, pair_mask), self.dz, 1)
        pair = pair + self._row_dropout(self.tri_mul_in(pair, pair_mask), self.dz, 1)
        pair = pair + self._row_dropout(self.tri_att_start(pair, pair_mask), self.dz, 1)
        pair = pair + self._row_dropout(self.tri_att_end(pair, pair_mask), self.dz, 2)
        return pair + self.pair_transition(pair)

    def forward(self, msa, pair, msa_mask, pair_mask):
        bp_size = scg.get_bp_world_size()
        if bp_size == 1:
            msa_c, mask_c = self._msa_branch(msa, pair, msa_mask)
            pair = pair + self.outer(msa_c, mask_c)
            msa = dap.col_to_row(msa_c)
            return msa, self._pair_branch(pair, pair_mask)
        # branch parallel: the outer-product-mean output of the *previous* MSA state feeds the pair branch
        rank = scg.get_bp_rank()
        msa_c0, mask_c0 = dap.row_to_col(msa), dap.row_to_col(msa_mask.unsqueeze(-1)).squeeze(-1)
        if rank == 0:
            new_msa_c, _ = self._msa_branch(msa, pair, msa_mask)
            new_msa, new_pair = dap.col_to_row(new_msa_c), pair
        else:
            new_pair = self._pair_branch(pair + self.outer(msa_c0, mask_c0), pair_mask)
            new_msa = msa
        return bp.sync_evoformer_results(new_msa, new_pair)


class EmbeddingsAndEvoformer(nn.Module):
    """Input embeddings (target / MSA features, relative positions, optional recycling) + N Evoformer blocks + single repr."""

    def __init__(self, msa_feat_dim=49, target_feat_dim=22, c_m=256, c_z=128, c_s=384, num_blocks=48, max_relative_feature=32,
                 msa_heads=8, pair_heads=4, use_recompute=False):
        super().__init__()
        self.preprocess_1d, self.preprocess_msa = nn.Linear(target_feat_dim, c_m), nn.Linear(msa_feat_dim, c_m)
        self.left_single, self.right_single = nn.Linear(target_feat_dim, c_z), nn.Linear(target_feat_dim, c_z)
        self.max_rel = max_relative_feature
        self.pair_relpos = nn.Linear(2 * max_relative_feature + 1, c_z)
        self.prev_pos_linear = nn.Linear(15, c_z)
        self.prev_msa_ln, self.prev_pair_ln = nn.LayerNorm(c_m), nn.LayerNorm(c_z)
        self.blocks = nn.ModuleList([EvoformerIteration(c_m, c_z, msa_heads, pair_heads) for _ in range(num_blocks)])
        self.single = nn.Linear(c_m, c_s)
        self.use_recompute = use_recompute

    def forward(self, batch, prev=None):
        tf, mf = batch["target_feat"], batch["msa_feat"]
        msa = self.preprocess_1d(tf)[:, None] + self.preprocess_msa(mf)
        pair = self.left_single(tf)[:, :, None] + self.right_single(tf)[:, None]
        res = batch["residue_index"]
        off = (res[:, :, None] - res[:, None, :]).clamp(-self.max_rel, self.max_rel) + self.max_rel
        pair = pair + self.pair_relpos(F.one_hot(off, 2 * self.max_rel + 1).to(pair.dtype))
        if prev is not None:
            msa = torch.cat([msa[:, :1] + self.prev_msa_ln(prev["prev_msa_first_row"])[:, None], msa[:, 1:]], 1)
            pair = pair + self.prev_pair_ln(prev["prev_pair"])
        msa_mask = batch.get("msa_mask", torch.ones(msa.shape[:3], device=msa.device, dtype=msa.dtype))
        seq_mask = batch.get("seq_mask", torch.ones(tf.shape[:2], device=msa.device, dtype=msa.dtype))
        pair_mask = seq_mask[:, :, None] * seq_mask[:, None, :]
        # enter the DAP layout: MSA sharded by sequences, pair by rows
        msa, msa_mask = dap.scatter(msa, 1), dap.scatter(msa_mask, 1)
        pair, pair_mask = dap.scatter(pair, 1), dap.scatter(pair_mask, 1)
        for blk in self.blocks:
            if self.use_recompute and self.training:
                from ...parallel.recompute import recompute

                msa, pair = recompute(blk, msa, pair, msa_mask, pair_mask)
            else:
                msa, pair = blk(msa, pair, msa_mask, pair_mask)
        msa, pair = dap.gather(msa, 1), dap.gather(pair, 1)
        return {"single": self.single(msa[:, 0]), "pair": pair, "msa": msa, "msa_first_row": msa[:, 0]}


DistEmbeddingsAndEvoformer = EmbeddingsAndEvoformer

"""Attention blocks of the folding trunk (reference ppfleetx/models/protein_folding/attentions.py:35-729): gated multi-head attention with
additive biases (the reference's ``fused_gate_attention``, :126-142), MSA row attention with pair bias, MSA column attention, global
(column-averaged) attention for the extra-MSA stack, and the two triangle updates of the pair representation.

Under DAP the MSA / pair activations arrive sharded (see ``evoformer.py``); the blocks gather what they need through
``distributed/protein_folding/dap.py``.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...distributed.protein_folding import dap
from ...ops.evoformer_attention import evoformer_attention

Attention = None  # set below: the reference's name for the gated attention primitive


class GatedAttention(nn.Module):
    """Multi-head attention with sigmoid output gating and an optional additive bias ``[b, h, q, k]`` (per leading group)."""

    def __init__(self, q_dim, kv_dim, num_head, out_dim, gating=True, key_dim=None):
        super().__init__()
        key_dim = key_dim or q_dim                      # total projection width (all heads)
        self.h, self.d = num_head, key_dim // num_head
        self.q, self.k, self.v = nn.Linear(q_dim, key_dim, bias=False), nn.Linear(kv_dim, key_dim, bias=False), nn.Linear(kv_dim, key_dim, bias=False)
        self.o = nn.Linear(key_dim, out_dim)
        self.g = nn.Linear(q_dim, key_dim) if gating else None
        if self.g is not None:
            nn.init.zeros_(self.g.weight); nn.init.ones_(self.g.bias)
        nn.init.zeros_(self.o.weight); nn.init.zeros_(self.o.bias)

    def forward(self, q_data, m_data, bias=None, nonbatched_bias=None):
        # q_data: [b, g, q, c]; m_data: [b, g, k, c]; bias: [b, g, 1, 1, k]; nonbatched_bias: [b, 1, h, q, k]
        b, g, nq, _ = q_data.shape
        nk = m_data.shape[2]
        q = self.q(q_data).view(b * g, nq, self.h, self.d)
        k = self.k(m_data).view(b * g, nk, self.h, self.d)
        v = self.v(m_data).view(b * g, nk, self.h, self.d)
        gate = self.g(q_data).view(b * g, nq, self.h, self.d) if self.g is not None else None
        mask_bias = bias.expand(b, g, 1, 1, nk).reshape(b * g, nk) if bias is not None else None          # broadcastable over b / g
        pair_bias = nonbatched_bias.expand(b, 1, self.h, nq, nk).reshape(b, self.h, nq, nk) if nonbatched_bias is not None else None
        # one fused kernel on B200 (logits, biases, softmax, P V and the gate stay on chip); the same expression in PyTorch elsewhere
        o = evoformer_attention(q, k, v, mask_bias, pair_bias, gate, groups_per_pair=g)
        return self.o(o.reshape(b, g, nq, self.h * self.d))


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
        z = dap.gather_full(z, axis=1)                             # full [b, R, R, h] on every rank
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


class GlobalAttention(nn.Module):
    """Global column-wise self-attention (Jumper et al. Suppl. Alg. 19): one mean query per column, keys / values shared by all heads, so
    the cost is linear in the number of sequences — what makes the ~1-5k-row extra-MSA stack affordable."""

    def __init__(self, q_dim, kv_dim, num_head, out_dim, gating=True, key_dim=None):
        super().__init__()
        key_dim = key_dim or q_dim
        self.h, self.d = num_head, key_dim // num_head
        self.q = nn.Linear(q_dim, key_dim, bias=False)
        self.k, self.v = nn.Linear(kv_dim, self.d, bias=False), nn.Linear(kv_dim, self.d, bias=False)
        self.o = nn.Linear(key_dim, out_dim)
        self.g = nn.Linear(q_dim, key_dim) if gating else None
        if self.g is not None:
            nn.init.zeros_(self.g.weight); nn.init.ones_(self.g.bias)
        nn.init.zeros_(self.o.weight); nn.init.zeros_(self.o.bias)

    def forward(self, q_data, m_data, q_mask):
        # q_data / m_data [b, g, n, c]; q_mask [b, g, n, 1]
        b, g, n, _ = q_data.shape
        q_mask = q_mask.to(q_data.dtype)
        q_avg = (q_data * q_mask).sum(2) / (q_mask.sum(2) + 1e-10)                              # [b, g, c]
        q = self.q(q_avg).view(b, g, self.h, self.d) * self.d ** -0.5
        k, v = self.k(m_data), self.v(m_data)                                                   # [b, g, n, d]
        logits = torch.einsum("bghd,bgnd->bghn", q, k) + (1e9 * (q_mask.squeeze(-1) - 1.0))[:, :, None, :]
        avg = torch.einsum("bghn,bgnd->bghd", torch.softmax(logits.float(), -1).to(v.dtype), v)  # [b, g, h, d]
        if self.g is not None:
            gate = torch.sigmoid(self.g(q_data)).view(b, g, n, self.h, self.d)
            out = (gate * avg[:, :, None]).reshape(b, g, n, self.h * self.d)
        else:
            out = avg.reshape(b, g, 1, self.h * self.d).expand(b, g, n, self.h * self.d)
        return self.o(out)


class MSAColumnGlobalAttention(nn.Module):
    def __init__(self, c_m, num_head=8):
        super().__init__()
        self.ln = nn.LayerNorm(c_m)
        self.attn = GlobalAttention(c_m, c_m, num_head, c_m)

    def forward(self, msa, msa_mask):
        m = self.ln(msa).transpose(1, 2)                          # [b, R(/n), S, c]
        mask = msa_mask.transpose(1, 2).unsqueeze(-1)
        return self.attn(m, m, mask).transpose(1, 2)


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
            b_full = dap.gather_full(b_, axis=1)
            x = torch.einsum("bikc,bjkc->bijc", a, b_full)
        else:
            # incoming: out[i,j] = sum_k a[k,i] b[k,j]: contraction runs over the sharded axis -> work on column shards
            a_c, b_c = dap.row_to_col(a), dap.row_to_col(b_)         # [b, R, R/n, c]
            b_full = dap.gather_full(b_c, axis=2)
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
        nb = dap.gather_full(self.bias(z), axis=1).permute(0, 3, 1, 2).unsqueeze(1)
        bias = (1e9 * (pair_mask - 1.0))[:, :, None, None, :]
        out = self.attn(z, z, bias, nb)
        if not self.starting:
            out = dap.col_to_row(out.transpose(1, 2))
        return out


Attention = GatedAttention

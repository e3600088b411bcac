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

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...distributed.protein_folding import bp, dap
from ...distributed.protein_folding.scg import scg


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
    """One Evoformer block.  ``outer_position`` places the outer-product-mean MSA -> pair update (reference evoformer.py:179-420):

    * ``"origin"`` — AlphaFold2 order: MSA branch, ``pair += OPM(new msa)``, pair branch.  Sequential, so no branch parallelism.
    * ``"end"``    — MSA branch and pair branch both start from the block's inputs, ``pair = pair_branch(pair) + OPM(new msa)``.  The two
      branches are independent, which is what branch parallelism (bp = 2) runs on two ranks; the unsharded model computes the same function.
    """

    def __init__(self, c_m=256, c_z=128, msa_heads=8, pair_heads=4, dropout_msa=0.15, dropout_pair=0.25, is_extra_msa=False, outer_position="origin"):
        super().__init__()
        assert outer_position in ("origin", "end"), outer_position
        self.msa_row = MSARowAttentionWithPairBias(c_m, c_z, msa_heads)
        self.is_extra_msa = is_extra_msa
        self.msa_col = MSAColumnGlobalAttention(c_m, msa_heads) if is_extra_msa else MSAColumnAttention(c_m, msa_heads)
        self.msa_transition = Transition(c_m)
        self.outer = OuterProductMean(c_m, c_z)
        self.tri_mul_out, self.tri_mul_in = TriangleMultiplication(c_z, outgoing=True), TriangleMultiplication(c_z, outgoing=False)
        self.tri_att_start, self.tri_att_end = TriangleAttention(c_z, pair_heads, True), TriangleAttention(c_z, pair_heads, False)
        self.pair_transition = Transition(c_z)
        self.dm, self.dz = dropout_msa, dropout_pair
        self.outer_position = outer_position

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
            if self.outer_position == "origin":
                new_pair = self._pair_branch(pair + self.outer(msa_c, mask_c), pair_mask)
            else:
                new_pair = self._pair_branch(pair, pair_mask) + self.outer(msa_c, mask_c)
            return dap.col_to_row(msa_c), new_pair
        assert bp_size == 2 and self.outer_position == "end", "branch parallelism needs bp_degree 2 and outer_position='end'"
        if scg.get_bp_rank() == 0:
            msa_c, mask_c = self._msa_branch(msa, pair, msa_mask)
            new_msa, outer, new_pair = dap.col_to_row(msa_c), self.outer(msa_c, mask_c), pair
        else:
            new_msa, outer, new_pair = msa, torch.zeros_like(pair), self._pair_branch(pair, pair_mask)
        return bp.sync_evoformer_results(new_msa, new_pair, outer)


class EmbeddingsAndEvoformer(nn.Module):
    """Input embeddings (target / MSA features, relative positions, recycling, templates, extra-MSA stack) + N Evoformer blocks + single
    representation (reference evoformer.py:532-996; Jumper et al. Suppl. Alg. 2 lines 5-18).

    Optional stages switch on with their inputs: the extra-MSA stack needs ``extra_msa_blocks > 0`` and ``batch['extra_msa']``; templates need
    ``template={'enabled': True, ...}`` and the ``template_*`` features; position recycling needs ``prev['prev_pos']`` + ``batch['aatype']``."""

    def __init__(self, msa_feat_dim=49, target_feat_dim=22, c_m=256, c_z=128, c_s=384, num_blocks=48, max_relative_feature=32,
                 msa_heads=8, pair_heads=4, use_recompute=False, extra_msa_channel=64, extra_msa_blocks=0, template=None,
                 prev_pos_bins=(15, 3.25, 20.75), outer_product_mean_position="origin"):
        super().__init__()
        self.preprocess_1d, self.preprocess_msa = nn.Linear(target_feat_dim, c_m), nn.Linear(msa_feat_dim, c_m)
        self.left_single, self.right_single = nn.Linear(target_feat_dim, c_z), nn.Linear(target_feat_dim, c_z)
        self.max_rel = max_relative_feature
        self.pair_relpos = nn.Linear(2 * max_relative_feature + 1, c_z)
        self.prev_pos_bins = tuple(prev_pos_bins)
        self.prev_pos_linear = nn.Linear(self.prev_pos_bins[0], c_z)
        self.prev_msa_ln, self.prev_pair_ln = nn.LayerNorm(c_m), nn.LayerNorm(c_z)
        # extra-MSA stack: 23 residue classes + has_deletion + deletion_value -> c_e; blocks use global column attention
        self.extra_msa_activations = nn.Linear(25, extra_msa_channel) if extra_msa_blocks else None
        pos = outer_product_mean_position
        self.extra_msa_stack = nn.ModuleList([EvoformerIteration(extra_msa_channel, c_z, msa_heads, pair_heads, is_extra_msa=True, outer_position=pos)
                                              for _ in range(extra_msa_blocks)])
        t = dict(template or {})
        self.template_enabled = bool(t.pop("enabled", False))
        self.embed_torsion_angles = bool(t.pop("embed_torsion_angles", False)) and self.template_enabled
        if self.template_enabled:
            from .template import TemplateEmbedding

            self.template_embedding = TemplateEmbedding(c_z=c_z, use_recompute=use_recompute, **t)
        if self.embed_torsion_angles:
            self.template_single_embedding, self.template_projection = nn.Linear(57, c_m), nn.Linear(c_m, c_m)
        self.blocks = nn.ModuleList([EvoformerIteration(c_m, c_z, msa_heads, pair_heads, outer_position=pos) for _ in range(num_blocks)])
        self.single = nn.Linear(c_m, c_s)
        self.use_recompute = use_recompute

    # -- gradient synchronisation contract (the reference flags optimizer param groups with ``dap`` / ``bp``, dap.py:400-425, bp.py:126-152)
    def dap_parameters(self):
        """Parameters applied to DAP-*sharded* activations: each rank holds a partial gradient, summed over the dap group.  Everything else
        sees replicated activations under DAP and already has the full gradient on every rank."""
        mods = [self.blocks, self.extra_msa_stack]
        if self.template_enabled:
            mods.append(self.template_embedding.single_template_embedding.template_pair_stack)
        return [p for m in mods for p in m.parameters()]

    def bp_parameters(self):
        """Parameters up to the trunk's exit: under branch parallelism each rank back-propagates only its branch's share into them, summed
        over the bp group.  The single-representation head runs after the exit on every rank and is excluded."""
        head = {id(p) for p in self.single.parameters()}
        return [p for p in self.parameters() if id(p) not in head]

    def sync_gradients(self):
        """Call after ``backward`` (before the data-parallel reduction / optimizer step)."""
        for group, params in ((dap, self.dap_parameters()), (bp, self.bp_parameters())):
            live = [p for p in params if p.requires_grad]
            for p in live:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            group.grad_sync(live)

    @staticmethod
    def pseudo_beta(aatype, all_atom_positions, all_atom_masks=None):
        """CB position (CA for glycine) and, given masks, its validity."""
        from . import residue_constants as rc

        is_gly = aatype == rc.restype_order["G"]
        ca, cb = rc.atom_order["CA"], rc.atom_order["CB"]
        pos = torch.where(is_gly.unsqueeze(-1), all_atom_positions[..., ca, :], all_atom_positions[..., cb, :])
        if all_atom_masks is None:
            return pos
        return pos, torch.where(is_gly, all_atom_masks[..., ca], all_atom_masks[..., cb])

    def _run_stack(self, blocks, msa, pair, msa_mask, pair_mask):
        for blk in blocks:
            if self.use_recompute and self.training:
                from ...parallel.recompute import recompute

                msa, pair = recompute(blk, msa, pair, msa_mask, pair_mask)
            else:
                msa, pair = blk(msa, pair, msa_mask, pair_mask)
        return msa, pair

    def forward(self, batch, prev=None):
        from .common import dgram_from_positions

        tf, mf = batch["target_feat"], batch["msa_feat"]
        msa = self.preprocess_1d(tf)[:, None] + self.preprocess_msa(mf)
        pair = self.left_single(tf)[:, :, None] + self.right_single(tf)[:, None]
        msa_mask = batch.get("msa_mask", torch.ones(msa.shape[:3], device=msa.device, dtype=msa.dtype))
        seq_mask = batch.get("seq_mask", torch.ones(tf.shape[:2], device=msa.device, dtype=msa.dtype))
        mask_2d = seq_mask[:, :, None] * seq_mask[:, None, :]
        if prev is not None:
            if "prev_pos" in prev and "aatype" in batch:
                pb = self.pseudo_beta(batch["aatype"], prev["prev_pos"])
                pair = pair + self.prev_pos_linear(dgram_from_positions(pb.float(), *self.prev_pos_bins).to(pair.dtype))
            if "prev_msa_first_row" in prev:
                msa = torch.cat([msa[:, :1] + self.prev_msa_ln(prev["prev_msa_first_row"])[:, None], msa[:, 1:]], 1)
            if "prev_pair" in prev:
                pair = pair + self.prev_pair_ln(prev["prev_pair"])
        res = batch["residue_index"]
        off = (res[:, :, None] - res[:, None, :]).clamp(-self.max_rel, self.max_rel) + self.max_rel
        pair = pair + self.pair_relpos(F.one_hot(off, 2 * self.max_rel + 1).to(pair.dtype))
        if self.template_enabled and "template_mask" in batch:
            tb = {k: v for k, v in batch.items() if k.startswith("template_")}
            pair = pair + self.template_embedding(pair, tb, mask_2d)
        if self.extra_msa_activations is not None and "extra_msa" in batch:
            feat = torch.cat([F.one_hot(batch["extra_msa"].long(), 23).to(pair.dtype), batch["extra_has_deletion"].unsqueeze(-1).to(pair.dtype),
                              batch["extra_deletion_value"].unsqueeze(-1).to(pair.dtype)], -1)
            e_msa = self.extra_msa_activations(feat)
            e_mask = batch.get("extra_msa_mask", torch.ones(e_msa.shape[:3], device=e_msa.device, dtype=e_msa.dtype))
            e_msa, e_mask_s = dap.scatter(e_msa, 1), dap.scatter(e_mask, 1)
            pair_s, pm_s = dap.scatter(pair, 1), dap.scatter(mask_2d, 1)
            _, pair_s = self._run_stack(self.extra_msa_stack, e_msa, pair_s, e_mask_s, pm_s)
            pair = dap.gather(pair_s, 1)
        if self.embed_torsion_angles and "template_aatype" in batch:
            from . import all_atom

            ret = all_atom.atom37_to_torsion_angles(batch["template_aatype"], batch["template_all_atom_positions"].float(),
                                                    batch["template_all_atom_masks"], placeholder_for_undefined=True)
            ret = {k: v.to(msa.dtype) for k, v in ret.items()}
            tfeat = torch.cat([F.one_hot(batch["template_aatype"].long(), 22).to(msa.dtype), ret["torsion_angles_sin_cos"].flatten(-2),
                               ret["alt_torsion_angles_sin_cos"].flatten(-2), ret["torsion_angles_mask"]], -1)          # 22 + 14 + 14 + 7 = 57
            tact = self.template_projection(F.relu(self.template_single_embedding(tfeat)))
            msa = torch.cat([msa, tact], 1)
            msa_mask = torch.cat([msa_mask, ret["torsion_angles_mask"][..., 2].to(msa_mask.dtype)], 1)
        # (branch parallelism needs no gradient broadcast here: each bp rank back-propagates its branch's share into the embeddings and
        # ``bp.grad_sync`` sums the replicated-parameter gradients)
        # enter the DAP layout: MSA sharded by sequences, pair by rows
        msa, msa_mask = dap.scatter(msa, 1), dap.scatter(msa_mask, 1)
        pair, pair_mask = dap.scatter(pair, 1), dap.scatter(mask_2d, 1)
        msa, pair = self._run_stack(self.blocks, msa, pair, msa_mask, pair_mask)
        msa, pair = dap.gather(msa, 1), dap.gather(pair, 1)
        msa, pair = bp.replicated_exit(msa), bp.replicated_exit(pair)     # heads / loss below run on every bp rank
        n_seq = mf.shape[1]
        return {"single": self.single(msa[:, 0]), "pair": pair, "msa": msa[:, :n_seq], "msa_first_row": msa[:, 0]}


DistEmbeddingsAndEvoformer = EmbeddingsAndEvoformer

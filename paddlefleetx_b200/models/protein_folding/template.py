"""Template embedding (reference ppfleetx/models/protein_folding/template.py:36-368; Jumper et al. Suppl. Alg. 2 lines 9-13, Alg. 16-17).

Each template's 2-D features — distogram of pseudo-beta atoms, residue types of both partners, the unit vector from every backbone frame
to every other CA, and validity masks (88 channels) — are embedded into ``c_t`` channels and refined by a small stack of triangle
updates (``TemplatePair``).  The query pair representation then attends over the templates point-wise (``TemplateEmbedding``) and the
result is added to the pair activations.

B200 note: the per-template stacks are independent, so they run as one batch of ``b * T`` problems (a single pass over larger GEMMs and
attention calls) instead of the reference's Python loop over templates.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...distributed.protein_folding import dap
from . import quat_affine
from . import residue_constants as rc
from .common import dgram_from_positions, no_autocast
from .evoformer import GatedAttention, Transition, TriangleAttention, TriangleMultiplication


class TemplatePairBlock(nn.Module):
    def __init__(self, c_t: int, num_head: int, dropout: float, transition_factor: int = 2, c_mul: int = 64):
        super().__init__()
        self.tri_att_start, self.tri_att_end = TriangleAttention(c_t, num_head, True), TriangleAttention(c_t, num_head, False)
        self.tri_mul_out, self.tri_mul_in = TriangleMultiplication(c_t, c_mul, outgoing=True), TriangleMultiplication(c_t, c_mul, outgoing=False)
        self.pair_transition = Transition(c_t, transition_factor)
        self.p = dropout

    def _drop(self, x, dim):
        if not self.training or self.p == 0:
            return x
        shape = list(x.shape)
        shape[dim] = 1
        return x * torch.bernoulli(torch.full(shape, 1 - self.p, device=x.device, dtype=x.dtype)) / (1 - self.p)

    def forward(self, act, mask):
        act = act + self._drop(self.tri_att_start(act, mask), 1)
        act = act + self._drop(self.tri_att_end(act, mask), 2)
        act = act + self._drop(self.tri_mul_out(act, mask), 1)
        act = act + self._drop(self.tri_mul_in(act, mask), 1)
        return act + self.pair_transition(act)


class TemplatePair(nn.Module):
    """Stack of ``num_block`` triangle-update blocks over one template's pair embedding (Suppl. Alg. 16 "TemplatePairStack")."""

    def __init__(self, c_t: int = 64, num_block: int = 2, num_head: int = 4, dropout: float = 0.25, use_recompute: bool = False):
        super().__init__()
        self.blocks = nn.ModuleList([TemplatePairBlock(c_t, num_head, dropout) for _ in range(num_block)])
        self.use_recompute = use_recompute

    def forward(self, act, mask):
        for blk in self.blocks:
            if self.use_recompute and self.training:
                from ...parallel.recompute import recompute

                act = recompute(blk, act, mask)
            else:
                act = blk(act, mask)
        return act


@no_autocast
def template_pair_features(batch, mask_2d, dtype, dgram_bins: int = 39, min_bin: float = 3.25, max_bin: float = 50.75,
                           use_template_unit_vector: bool = False) -> torch.Tensor:
    """-> ``[B, R, R, 88]`` for ``B`` templates given ``template_aatype [B, R]``, ``template_pseudo_beta [B, R, 3]`` (+ mask),
    ``template_all_atom_positions [B, R, 37, 3]`` (+ masks)."""
    out_dtype, dtype = dtype, torch.float32          # geometry (frames, distances) stays in fp32 whatever the trunk's compute dtype
    pb_mask = batch["template_pseudo_beta_mask"].to(dtype)
    pb_mask_2d = pb_mask[:, :, None] * pb_mask[:, None, :]
    dgram = dgram_from_positions(batch["template_pseudo_beta"].to(dtype), dgram_bins, min_bin, max_bin)
    aatype = F.one_hot(batch["template_aatype"].long(), 22).to(dtype)
    R = aatype.shape[1]
    feats = [dgram, pb_mask_2d.unsqueeze(-1), aatype[:, None, :, :].expand(-1, R, -1, -1), aatype[:, :, None, :].expand(-1, -1, R, -1)]
    n, ca, c = (rc.atom_order[a] for a in ("N", "CA", "C"))
    pos, amask = batch["template_all_atom_positions"].to(dtype), batch["template_all_atom_masks"].to(dtype)
    rot, trans = quat_affine.make_transform_from_reference(pos[:, :, n], pos[:, :, ca], pos[:, :, c])
    frames = quat_affine.QuatAffine(None, trans, rotation=rot)
    vec = frames.invert_point(trans[:, None, :, :], extra_dims=1)                         # [B, R(frame i), R(point j), 3]
    unit = vec * torch.rsqrt(1e-6 + (vec * vec).sum(-1, keepdim=True))
    bb_mask = amask[:, :, n] * amask[:, :, ca] * amask[:, :, c]
    bb_mask_2d = bb_mask[:, :, None] * bb_mask[:, None, :]
    unit = unit * bb_mask_2d.unsqueeze(-1)
    if not use_template_unit_vector:
        unit = torch.zeros_like(unit)
    feats += [unit, bb_mask_2d.unsqueeze(-1)]
    return (torch.cat(feats, dim=-1) * bb_mask_2d.unsqueeze(-1)).to(out_dtype)


class SingleTemplateEmbedding(nn.Module):
    """Features of a batch of templates -> ``[B, R, R, c_t]`` (Suppl. Alg. 2 lines 9-11)."""

    FEATURE_DIM = 39 + 1 + 22 + 22 + 3 + 1

    def __init__(self, c_t: int = 64, num_block: int = 2, num_head: int = 4, dropout: float = 0.25, dgram_bins: int = 39, min_bin: float = 3.25,
                 max_bin: float = 50.75, use_template_unit_vector: bool = False, use_recompute: bool = False):
        super().__init__()
        self.dgram = dict(dgram_bins=dgram_bins, min_bin=min_bin, max_bin=max_bin, use_template_unit_vector=use_template_unit_vector)
        self.embedding2d = nn.Linear(self.FEATURE_DIM - 39 + dgram_bins, c_t)
        nn.init.kaiming_normal_(self.embedding2d.weight, nonlinearity="relu")
        self.template_pair_stack = TemplatePair(c_t, num_block, num_head, dropout, use_recompute)
        self.output_layer_norm = nn.LayerNorm(c_t)

    def forward(self, batch, mask_2d, dtype):
        act = self.embedding2d(template_pair_features(batch, mask_2d, dtype, **self.dgram))
        # DAP layout for the triangle updates: rows sharded
        act, m = dap.scatter(act, 1), dap.scatter(mask_2d, 1)
        act = self.template_pair_stack(act, m)
        return self.output_layer_norm(dap.gather(act, 1))


class TemplateEmbedding(nn.Module):
    """All templates -> one additive update of the pair representation (Suppl. Alg. 17 "TemplatePointwiseAttention")."""

    def __init__(self, c_z: int = 128, c_t: int = 64, num_block: int = 2, num_head: int = 4, attn_key_dim: int = 64, dropout: float = 0.25,
                 use_template_unit_vector: bool = False, use_recompute: bool = False, **single_kw):
        super().__init__()
        self.single_template_embedding = SingleTemplateEmbedding(c_t, num_block, num_head, dropout, use_template_unit_vector=use_template_unit_vector,
                                                                 use_recompute=use_recompute, **single_kw)
        self.attention = GatedAttention(c_z, c_t, num_head, c_z, gating=False, key_dim=attn_key_dim)

    def forward(self, query_embedding, template_batch, mask_2d):
        """query_embedding ``[b, R, R, c_z]``; template_batch tensors ``[b, T, ...]``; mask_2d ``[b, R, R]`` -> ``[b, R, R, c_z]``."""
        b, T = template_batch["template_mask"].shape
        R, dtype = query_embedding.shape[1], query_embedding.dtype
        flat = {k: v.flatten(0, 1) for k, v in template_batch.items() if k != "template_mask"}
        m2 = mask_2d.to(dtype)[:, None].expand(-1, T, -1, -1).flatten(0, 1)
        rep = self.single_template_embedding(flat, m2, dtype).unflatten(0, (b, T))          # [b, T, R, R, c_t]
        q = query_embedding.reshape(b, R * R, 1, -1)
        kv = rep.permute(0, 2, 3, 1, 4).reshape(b, R * R, T, -1)
        tmask = template_batch["template_mask"].to(dtype)
        bias = (1e9 * (tmask - 1.0))[:, None, None, None, :]                                  # [b, 1, 1, 1, T]
        out = self.attention(q, kv, bias).reshape(b, R, R, -1)
        return out * (tmask.sum(-1) > 0).to(dtype)[:, None, None, None]

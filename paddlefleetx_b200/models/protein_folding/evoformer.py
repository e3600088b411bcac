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


import torch
import torch.nn as nn
import torch.nn.functional as F

from ...distributed.protein_folding import bp, dap
from ...distributed.protein_folding.scg import scg
from .attentions import (GatedAttention, GlobalAttention, MSAColumnAttention, MSAColumnGlobalAttention, MSARowAttentionWithPairBias,  # noqa: F401
                         TriangleAttention, TriangleMultiplication)
from .outer_product_mean import OuterProductMean  # noqa: F401


class Transition(nn.Module):
    def __init__(self, c, factor=4):
        super().__init__()
        self.ln = nn.LayerNorm(c)
        self.fc1, self.fc2 = nn.Linear(c, c * factor), nn.Linear(c * factor, c)
        nn.init.zeros_(self.fc2.weight); nn.init.zeros_(self.fc2.bias)

    def forward(self, x, mask=None):
        return self.fc2(F.relu(self.fc1(self.ln(x))))


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
        """Features -> ``{single, pair, msa}`` representations: target / MSA embedding, relative-position and recycling (``prev``) terms, template
        pair stack + template attention, the extra-MSA stack, the Evoformer blocks, template torsion-angle rows and the single projection
        (reference evoformer.py:504-760)."""
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

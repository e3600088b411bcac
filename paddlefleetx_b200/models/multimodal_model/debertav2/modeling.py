"""DeBERTa-V2 encoder (alternative Imagen text tower) — reference models/language_model/debertav2/modeling.py:31-1323.

Disentangled attention: every score is content-to-content plus content-to-position plus position-to-content, the position terms indexed by
log-bucketed relative distance into ONE table of relative embeddings (layer-normed, shared by all layers, projected with the layer's own key /
query projections when ``share_att_key``); a convolution branch after the first layer; masked softmax that zeroes fully masked rows.

The module tree and state-dict keys are the reference's (and the upstream checkpoints'): ``embeddings.{word_embeddings, LayerNorm, ...}``,
``encoder.layer.{i}.attention.self.{query_proj, key_proj, value_proj}``, ``...attention.output.{dense, LayerNorm}``, ``...intermediate.dense``,
``...output.{dense, LayerNorm}``, ``encoder.rel_embeddings``, ``encoder.LayerNorm``, ``encoder.conv.{conv, LayerNorm}`` — a converted
``debertav2.pd`` loads by name.
"""
from __future__ import annotations

import json
import math
import os
from collections.abc import Sequence
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ....ops import functional as OF
from ...language_model.ernie.model_outputs import ModelOutput
from ..t5.modeling import ACT2FN, _read_checkpoint


@dataclass
class BaseModelOutput(ModelOutput):
    """``last_hidden_state`` [b, s, h]; ``hidden_states`` (embedding output + one per layer) and ``attentions`` [b, heads, s, s] on request."""
    last_hidden_state: Optional[torch.Tensor] = None
    hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    attentions: Optional[Tuple[torch.Tensor, ...]] = None


# ----------------------------------------------------------------------------------------------------------------- masked softmax / dropout
class XSoftmax(torch.autograd.Function):
    """``XSoftmax.apply(input, mask, dim)``: softmax over the positions where ``mask`` is non-zero; masked positions get probability 0, also in
    rows that are masked entirely (a plain softmax would spread 1/n there).  Computed in fp32 (reference modeling.py:57-97)."""

    @staticmethod
    def forward(ctx, input, mask, dim):
        ctx.dim = dim
        off = ~mask.to(torch.bool)
        out = torch.softmax(input.float().masked_fill(off, torch.finfo(torch.float32).min), dim)
        out = out.masked_fill(off, 0.0)
        ctx.save_for_backward(out)
        ctx.in_dtype = input.dtype
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (out,) = ctx.saved_tensors
        g = grad_output.float()
        return (out * (g - (g * out).sum(ctx.dim, keepdim=True))).to(ctx.in_dtype), None, None


class DropoutContext:
    """Holds one dropout mask so that it can be re-used (``reuse_mask``) with a rescaled rate (``scale``) across calls."""

    def __init__(self):
        self.dropout = 0
        self.mask = None
        self.scale = 1
        self.reuse_mask = True


def get_mask(input, local_context):
    """``(mask, rate)``: True where the element is DROPPED.  ``local_context`` is a rate or a ``DropoutContext`` (whose stored mask is re-used)."""
    if isinstance(local_context, DropoutContext):
        dropout = local_context.dropout * local_context.scale
        mask = local_context.mask if local_context.reuse_mask else None
    else:
        dropout, mask = local_context, None
    if dropout > 0 and mask is None:
        mask = torch.rand_like(input, dtype=torch.float32) < dropout
    if isinstance(local_context, DropoutContext) and local_context.mask is None:
        local_context.mask = mask
    return mask, dropout


class XDropout(torch.autograd.Function):
    """Dropout by a boolean mask (no float mask tensor kept for backward)."""

    @staticmethod
    def forward(ctx, input, local_ctx):
        mask, dropout = get_mask(input, local_ctx)
        ctx.scale = 1.0 / (1 - dropout)
        if dropout > 0:
            ctx.save_for_backward(mask)
            return input.masked_fill(mask, 0) * ctx.scale
        return input

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.scale > 1:
            (mask,) = ctx.saved_tensors
            return grad_output.masked_fill(mask, 0) * ctx.scale, None
        return grad_output, None


class StableDropout(nn.Module):
    """Dropout module with an optional stack of contexts: after ``init_context()`` successive calls record their masks, and a second pass (after
    another ``init_context(reuse_mask=True, scale=...)``) replays them (reference modeling.py:146-193)."""

    def __init__(self, drop_prob):
        super().__init__()
        self.drop_prob = drop_prob
        self.count = 0
        self.context_stack = None

    def forward(self, x):
        if self.training and self.drop_prob > 0:
            return XDropout.apply(x, self.get_context())
        return x

    def clear_context(self):
        self.count = 0
        self.context_stack = None

    def init_context(self, reuse_mask=True, scale=1):
        if self.context_stack is None:
            self.context_stack = []
        self.count = 0
        for c in self.context_stack:
            c.reuse_mask, c.scale = reuse_mask, scale

    def get_context(self):
        if self.context_stack is None:
            return self.drop_prob
        if self.count >= len(self.context_stack):
            self.context_stack.append(DropoutContext())
        ctx = self.context_stack[self.count]
        ctx.dropout = self.drop_prob
        self.count += 1
        return ctx


# ----------------------------------------------------------------------------------------------------------------- relative positions
def make_log_bucket_position(relative_pos, bucket_size, max_position):
    """Signed distance -> bucket: distances below ``bucket_size / 2`` keep their value, larger ones are log-spaced up to ``max_position``."""
    sign = torch.sign(relative_pos)
    mid = bucket_size // 2
    abs_pos = torch.where((relative_pos < mid) & (relative_pos > -mid), torch.full_like(relative_pos, mid - 1), relative_pos.abs())
    log_pos = torch.ceil(torch.log(abs_pos.float() / mid) / math.log((max_position - 1) / mid) * (mid - 1)) + mid
    return torch.where(abs_pos <= mid, relative_pos.float(), log_pos * sign).long()


def build_relative_position(query_size, key_size, bucket_size=-1, max_position=-1, device=None):
    """``[1, query_size, key_size]`` of (bucketed) ``query position - key position``."""
    rel = torch.arange(query_size, device=device)[:, None] - torch.arange(key_size, device=device)[None, :]
    if bucket_size > 0 and max_position > 0:
        rel = make_log_bucket_position(rel, bucket_size, max_position)
    return rel.long().unsqueeze(0)


def c2p_dynamic_expand(c2p_pos, query_layer, relative_pos):
    return c2p_pos.expand(query_layer.shape[0], query_layer.shape[1], query_layer.shape[2], relative_pos.shape[-1])


def p2c_dynamic_expand(c2p_pos, query_layer, key_layer):
    return c2p_pos.expand(query_layer.shape[0], query_layer.shape[1], key_layer.shape[-2], key_layer.shape[-2])


def pos_dynamic_expand(pos_index, p2c_att, key_layer):
    return pos_index.expand(tuple(p2c_att.shape[:2]) + (pos_index.shape[-2], key_layer.shape[-2]))


# ----------------------------------------------------------------------------------------------------------------- layers
def _ln(hidden_size, eps, dtype, device):
    return nn.LayerNorm(hidden_size, eps, dtype=dtype, device=device)


class DisentangledSelfAttention(nn.Module):
    def __init__(self, hidden_size=1536, num_attention_heads=24, attention_head_size=None, share_att_key=False, pos_att_type=None,
                 relative_attention=False, position_buckets=-1, max_relative_positions=-1, max_position_embeddings=512, hidden_dropout_prob=0.0,
                 attention_probs_dropout_prob=0.0, dtype=None, device=None):
        super().__init__()
        if hidden_size % num_attention_heads != 0:
            raise ValueError(f"The hidden size ({hidden_size}) is not a multiple of the number of attention heads ({num_attention_heads})")
        self.num_attention_heads = num_attention_heads
        self.attention_head_size = attention_head_size if attention_head_size is not None else hidden_size // num_attention_heads
        self.all_head_size = num_attention_heads * self.attention_head_size
        kw = dict(dtype=dtype, device=device)
        self.query_proj, self.key_proj, self.value_proj = (nn.Linear(hidden_size, self.all_head_size, **kw) for _ in range(3))
        self.share_att_key = share_att_key
        self.pos_att_type = list(pos_att_type) if pos_att_type is not None else []
        self.relative_attention = relative_attention
        if relative_attention:
            self.position_buckets = position_buckets
            self.max_relative_positions = max_relative_positions if max_relative_positions >= 1 else max_position_embeddings
            self.pos_ebd_size = position_buckets if position_buckets > 0 else self.max_relative_positions
            self.pos_dropout = StableDropout(hidden_dropout_prob)
            if not share_att_key:
                if "c2p" in self.pos_att_type:
                    self.pos_key_proj = nn.Linear(hidden_size, self.all_head_size, **kw)
                if "p2c" in self.pos_att_type:
                    self.pos_query_proj = nn.Linear(hidden_size, self.all_head_size, **kw)
        self.dropout = StableDropout(attention_probs_dropout_prob)

    def transpose_for_scores(self, x, attention_heads):
        """[b, s, heads * d] -> [b * heads, s, d]."""
        b, s = x.shape[:2]
        return x.view(b, s, attention_heads, -1).transpose(1, 2).reshape(b * attention_heads, s, -1)

    def forward(self, hidden_states, attention_mask, output_attentions=False, query_states=None, relative_pos=None, rel_embeddings=None):
        """``attention_mask`` [b, 1, q, k] with 1 = may attend.  ``query_states`` (enhanced-mask-decoder passes) provides the queries while keys /
        values come from ``hidden_states``.  Returns the context [b, q, heads * d] (and the probabilities [b, heads, q, k])."""
        if query_states is None:
            query_states = hidden_states
        heads = self.num_attention_heads
        q = self.transpose_for_scores(OF.linear(query_states, self.query_proj.weight, self.query_proj.bias), heads)
        k = self.transpose_for_scores(OF.linear(hidden_states, self.key_proj.weight, self.key_proj.bias), heads)
        v = self.transpose_for_scores(OF.linear(hidden_states, self.value_proj.weight, self.value_proj.bias), heads)
        scale_factor = 1 + ("c2p" in self.pos_att_type) + ("p2c" in self.pos_att_type)
        scores = torch.bmm(q, k.transpose(1, 2)) / math.sqrt(q.shape[-1] * scale_factor)
        if self.relative_attention:
            scores = scores + self.disentangled_attention_bias(q, k, relative_pos, self.pos_dropout(rel_embeddings), scale_factor)
        scores = scores.view(-1, heads, scores.shape[-2], scores.shape[-1])
        probs = self.dropout(XSoftmax.apply(scores, attention_mask, -1).to(v.dtype))
        ctx = torch.bmm(probs.reshape(-1, probs.shape[-2], probs.shape[-1]), v)
        ctx = ctx.view(-1, heads, ctx.shape[-2], ctx.shape[-1]).transpose(1, 2)
        ctx = ctx.reshape(ctx.shape[0], ctx.shape[1], -1)
        return (ctx, probs) if output_attentions else ctx

    def disentangled_attention_bias(self, query_layer, key_layer, relative_pos, rel_embeddings, scale_factor):
        """content->position + position->content scores, [b * heads, q, k]."""
        q_len, k_len = query_layer.shape[-2], key_layer.shape[-2]
        if relative_pos is None:
            relative_pos = build_relative_position(q_len, k_len, self.position_buckets, self.max_relative_positions, query_layer.device)
        if relative_pos.dim() == 2:
            relative_pos = relative_pos[None, None]
        elif relative_pos.dim() == 3:
            relative_pos = relative_pos.unsqueeze(1)
        elif relative_pos.dim() != 4:
            raise ValueError(f"Relative position ids must be of dim 2 or 3 or 4. {relative_pos.dim()}")
        span = self.pos_ebd_size
        relative_pos = relative_pos.long().to(query_layer.device)
        rel = rel_embeddings[: span * 2].unsqueeze(0)
        heads, rep = self.num_attention_heads, query_layer.shape[0] // self.num_attention_heads
        key_proj = self.key_proj if self.share_att_key else getattr(self, "pos_key_proj", None)
        query_proj = self.query_proj if self.share_att_key else getattr(self, "pos_query_proj", None)
        score = 0
        if "c2p" in self.pos_att_type:
            pos_key = self.transpose_for_scores(OF.linear(rel, key_proj.weight, key_proj.bias), heads).repeat(rep, 1, 1)      # [b*h, 2span, d]
            c2p = torch.bmm(query_layer, pos_key.transpose(1, 2))                                                                # [b*h, q, 2span]
            idx = torch.clamp(relative_pos + span, 0, span * 2 - 1).squeeze(0).expand(query_layer.shape[0], q_len, relative_pos.shape[-1])
            score = score + torch.gather(c2p, -1, idx) / math.sqrt(pos_key.shape[-1] * scale_factor)
        if "p2c" in self.pos_att_type:
            pos_query = self.transpose_for_scores(OF.linear(rel, query_proj.weight, query_proj.bias), heads).repeat(rep, 1, 1)
            r_pos = relative_pos if k_len == q_len else build_relative_position(k_len, k_len, self.position_buckets, self.max_relative_positions,
                                                                                  query_layer.device)[None]
            idx = torch.clamp(-r_pos + span, 0, span * 2 - 1).squeeze(0).expand(query_layer.shape[0], k_len, k_len)
            p2c = torch.gather(torch.bmm(key_layer, pos_query.transpose(1, 2)), -1, idx).transpose(1, 2)
            if k_len != q_len:            # queries are the LAST q_len positions' worth of rows only when lengths agree; otherwise align by row index
                p2c = p2c[:, :q_len]
            score = score + p2c / math.sqrt(pos_query.shape[-1] * scale_factor)
        return score


class DebertaV2SelfOutput(nn.Module):
    def __init__(self, hidden_size=1536, layer_norm_eps=1e-7, hidden_dropout_prob=0.1, dtype=None, device=None):
        super().__init__()
        self.dense = nn.Linear(hidden_size, hidden_size, dtype=dtype, device=device)
        self.LayerNorm = _ln(hidden_size, layer_norm_eps, dtype, device)
        self.dropout = StableDropout(hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        return self.LayerNorm(self.dropout(OF.linear(hidden_states, self.dense.weight, self.dense.bias)) + input_tensor)


class DebertaV2Attention(nn.Module):
    def __init__(self, hidden_size=512, num_attention_heads=24, attention_head_size=64, share_att_key=True, pos_att_type=None,
                 relative_attention=True, position_buckets=-1, max_relative_positions=-1, max_position_embeddings=512, layer_norm_eps=1e-7,
                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, dtype=None, device=None):
        super().__init__()
        self.self = DisentangledSelfAttention(hidden_size, num_attention_heads, attention_head_size, share_att_key, pos_att_type, relative_attention,
                                              position_buckets, max_relative_positions, max_position_embeddings, hidden_dropout_prob,
                                              attention_probs_dropout_prob, dtype, device)
        self.output = DebertaV2SelfOutput(hidden_size, layer_norm_eps, hidden_dropout_prob, dtype, device)

    def forward(self, hidden_states, attention_mask, output_attentions=False, query_states=None, relative_pos=None, rel_embeddings=None):
        out = self.self(hidden_states, attention_mask, output_attentions, query_states=query_states, relative_pos=relative_pos,
                        rel_embeddings=rel_embeddings)
        att = None
        if output_attentions:
            out, att = out
        out = self.output(out, hidden_states if query_states is None else query_states)
        return (out, att) if output_attentions else out


class DebertaV2Intermediate(nn.Module):
    def __init__(self, hidden_size=1536, hidden_act="gelu", intermediate_size=6144, dtype=None, device=None):
        super().__init__()
        self.dense = nn.Linear(hidden_size, intermediate_size, dtype=dtype, device=device)
        self.intermediate_act_fn = ACT2FN[hidden_act] if isinstance(hidden_act, str) else hidden_act

    def forward(self, hidden_states):
        return self.intermediate_act_fn(OF.linear(hidden_states, self.dense.weight, self.dense.bias))


class DebertaV2Output(nn.Module):
    def __init__(self, hidden_size=512, intermediate_size=6144, layer_norm_eps=1e-7, hidden_dropout_prob=0.1, dtype=None, device=None):
        super().__init__()
        self.dense = nn.Linear(intermediate_size, hidden_size, dtype=dtype, device=device)
        self.LayerNorm = _ln(hidden_size, layer_norm_eps, dtype, device)
        self.dropout = StableDropout(hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        return self.LayerNorm(self.dropout(OF.linear(hidden_states, self.dense.weight, self.dense.bias)) + input_tensor)


class DebertaV2Layer(nn.Module):
    def __init__(self, hidden_size=512, hidden_act="gelu", intermediate_size=6144, num_attention_heads=24, attention_head_size=64, share_att_key=True,
                 pos_att_type=None, relative_attention=True, position_buckets=256, max_relative_positions=-1, max_position_embeddings=512,
                 layer_norm_eps=1e-7, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, dtype=None, device=None):
        super().__init__()
        self.attention = DebertaV2Attention(hidden_size, num_attention_heads, attention_head_size, share_att_key, pos_att_type, relative_attention,
                                            position_buckets, max_relative_positions, max_position_embeddings, layer_norm_eps, hidden_dropout_prob,
                                            attention_probs_dropout_prob, dtype, device)
        self.intermediate = DebertaV2Intermediate(hidden_size, hidden_act, intermediate_size, dtype, device)
        self.output = DebertaV2Output(hidden_size, intermediate_size, layer_norm_eps, hidden_dropout_prob, dtype, device)

    def forward(self, hidden_states, attention_mask, query_states=None, relative_pos=None, rel_embeddings=None, output_attentions=False):
        att_out = self.attention(hidden_states, attention_mask, output_attentions=output_attentions, query_states=query_states,
                                 relative_pos=relative_pos, rel_embeddings=rel_embeddings)
        att = None
        if output_attentions:
            att_out, att = att_out
        out = self.output(self.intermediate(att_out), att_out)
        return (out, att) if output_attentions else out


class ConvLayer(nn.Module):
    """Convolution over the sequence on the embedding output, added to the first layer's output (reference modeling.py:381-425)."""

    def __init__(self, hidden_size=512, conv_kernel_size=3, conv_groups=1, conv_act="tanh", layer_norm_eps=1e-7, hidden_dropout_prob=0.0, dtype=None,
                 device=None):
        super().__init__()
        self.conv_act = conv_act
        self.conv = nn.Conv1d(hidden_size, hidden_size, conv_kernel_size, padding=(conv_kernel_size - 1) // 2, groups=conv_groups, dtype=dtype,
                              device=device)
        self.LayerNorm = _ln(hidden_size, layer_norm_eps, dtype, device)
        self.dropout = StableDropout(hidden_dropout_prob)

    def forward(self, hidden_states, residual_states, input_mask):
        out = self.conv(hidden_states.transpose(1, 2)).transpose(1, 2)
        keep = None
        if input_mask is not None:
            keep = input_mask
            if keep.dim() == 4:
                keep = keep.squeeze(1).squeeze(1)
            keep = keep.to(torch.bool).unsqueeze(-1)
            out = out.masked_fill(~keep, 0.0)
        out = ACT2FN[self.conv_act](self.dropout(out))
        out = self.LayerNorm(residual_states + out)
        return out if keep is None else out * keep.to(out.dtype)


class DebertaV2Encoder(nn.Module):
    """Layer stack with the shared relative-embedding table."""

    def __init__(self, num_hidden_layers=48, num_attention_heads=24, attention_head_size=64, relative_attention=False, max_relative_positions=-1,
                 max_position_embeddings=512, position_buckets=256, hidden_size=1536, hidden_act="gelu", conv_act="gelu", intermediate_size=6144,
                 share_att_key=True, pos_att_type=None, norm_rel_ebd=None, conv_kernel_size=0, layer_norm_eps=1e-7, hidden_dropout_prob=0.1,
                 attention_probs_dropout_prob=0.1, dtype=None, device=None):
        super().__init__()
        self.layer = nn.ModuleList([DebertaV2Layer(hidden_size, hidden_act, intermediate_size, num_attention_heads, attention_head_size, share_att_key,
                                                   pos_att_type, relative_attention, position_buckets, max_relative_positions,
                                                   max_position_embeddings, layer_norm_eps, hidden_dropout_prob, attention_probs_dropout_prob, dtype,
                                                   device) for _ in range(num_hidden_layers)])
        self.relative_attention = relative_attention
        if relative_attention:
            self.max_relative_positions = max_relative_positions if max_relative_positions >= 1 else max_position_embeddings
            self.position_buckets = position_buckets
            size = (position_buckets if position_buckets > 0 else self.max_relative_positions) * 2
            self.rel_embeddings = nn.Embedding(size, hidden_size, dtype=dtype, device=device)
        self.norm_rel_ebd = [x.strip() for x in (norm_rel_ebd or "none").lower().split("|")]
        if "layer_norm" in self.norm_rel_ebd:
            self.LayerNorm = _ln(hidden_size, layer_norm_eps, dtype, device)
        self.conv = ConvLayer(hidden_size, conv_kernel_size, 1, conv_act, layer_norm_eps, hidden_dropout_prob, dtype, device) \
            if conv_kernel_size > 0 else None
        self.gradient_checkpointing = False

    def get_rel_embedding(self):
        rel = self.rel_embeddings.weight if self.relative_attention else None
        if rel is not None and "layer_norm" in self.norm_rel_ebd:
            rel = self.LayerNorm(rel)
        return rel

    def get_attention_mask(self, attention_mask):
        """[b, s] keep-mask -> [b, 1, s, s] (a pair attends iff both tokens are real); [b, q, k] -> [b, 1, q, k]."""
        if attention_mask.dim() <= 2:
            ext = attention_mask[:, None, None, :]
            attention_mask = (ext * ext.squeeze(-2).unsqueeze(-1)).to(torch.uint8)
        elif attention_mask.dim() == 3:
            attention_mask = attention_mask.unsqueeze(1)
        return attention_mask

    def get_rel_pos(self, hidden_states, query_states=None, relative_pos=None):
        if self.relative_attention and relative_pos is None:
            q = query_states.shape[-2] if query_states is not None else hidden_states.shape[-2]
            relative_pos = build_relative_position(q, hidden_states.shape[-2], self.position_buckets, self.max_relative_positions, hidden_states.device)
        return relative_pos

    def forward(self, hidden_states, attention_mask, output_hidden_states=True, output_attentions=False, query_states=None, relative_pos=None,
                return_dict=True):
        input_mask = attention_mask if attention_mask.dim() <= 2 else (attention_mask.sum(-2) > 0).to(torch.uint8)
        attention_mask = self.get_attention_mask(attention_mask)
        first = hidden_states[0] if isinstance(hidden_states, Sequence) else hidden_states
        relative_pos = self.get_rel_pos(first, query_states, relative_pos)
        all_hidden = () if output_hidden_states else None
        all_att = () if output_attentions else None
        next_kv, out = first, first
        rel = self.get_rel_embedding()
        for i, layer in enumerate(self.layer):
            if output_hidden_states:
                all_hidden = all_hidden + (out,)
            if self.gradient_checkpointing and self.training:
                from ....parallel.recompute import recompute

                out = recompute(layer, next_kv, attention_mask, query_states, relative_pos, rel, output_attentions)
            else:
                out = layer(next_kv, attention_mask, query_states=query_states, relative_pos=relative_pos, rel_embeddings=rel,
                            output_attentions=output_attentions)
            att = None
            if output_attentions:
                out, att = out
            if i == 0 and self.conv is not None:
                out = self.conv(first, out, input_mask)
            if query_states is not None:
                query_states = out
                if isinstance(hidden_states, Sequence):
                    next_kv = hidden_states[i + 1] if i + 1 < len(self.layer) else None
            else:
                next_kv = out
            if output_attentions:
                all_att = all_att + (att,)
        if output_hidden_states:
            all_hidden = all_hidden + (out,)
        if not return_dict:
            return tuple(v for v in (out, all_hidden, all_att) if v is not None)
        return BaseModelOutput(last_hidden_state=out, hidden_states=all_hidden, attentions=all_att)


class DebertaV2Embeddings(nn.Module):
    """Word (+ optional absolute position, token type) embeddings, optional projection to the hidden size, layer norm, padding zeroed."""

    def __init__(self, max_position_embeddings=512, position_biased_input=False, pad_token_id=0, hidden_size=1536, hidden_dropout_prob=0.1,
                 embedding_size=None, vocab_size=128100, type_vocab_size=0, layer_norm_eps=1e-7, dtype=None, device=None):
        super().__init__()
        kw = dict(dtype=dtype, device=device)
        self.embedding_size = hidden_size if embedding_size is None else embedding_size
        self.word_embeddings = nn.Embedding(vocab_size, self.embedding_size, padding_idx=pad_token_id, **kw)
        self.type_vocab_size, self.hidden_size, self.position_biased_input = type_vocab_size, hidden_size, position_biased_input
        self.position_embeddings = nn.Embedding(max_position_embeddings, self.embedding_size, **kw) if position_biased_input else None
        if type_vocab_size > 0:
            self.token_type_embeddings = nn.Embedding(type_vocab_size, self.embedding_size, **kw)
        if self.embedding_size != hidden_size:
            self.embed_proj = nn.Linear(self.embedding_size, hidden_size, **kw)
        self.LayerNorm = _ln(hidden_size, layer_norm_eps, dtype, device)
        self.dropout = StableDropout(hidden_dropout_prob)
        self.register_buffer("position_ids", torch.arange(max_position_embeddings, device=device).expand(1, -1), persistent=False)

    def forward(self, input_ids=None, token_type_ids=None, position_ids=None, mask=None, inputs_embeds=None):
        shape = input_ids.shape if input_ids is not None else inputs_embeds.shape[:-1]
        if inputs_embeds is None:
            inputs_embeds = self.word_embeddings(input_ids)
        emb = inputs_embeds
        if self.position_embeddings is not None:
            if position_ids is None:
                position_ids = self.position_ids[:, :shape[1]]
            emb = emb + self.position_embeddings(position_ids.long())
        if self.type_vocab_size > 0:
            if token_type_ids is None:
                token_type_ids = torch.zeros(tuple(shape), dtype=torch.long, device=emb.device)
            emb = emb + self.token_type_embeddings(token_type_ids)
        if self.embedding_size != self.hidden_size:
            emb = self.embed_proj(emb)
        emb = self.LayerNorm(emb)
        if mask is not None:
            if mask.dim() != emb.dim():
                if mask.dim() == 4:
                    mask = mask.squeeze(1).squeeze(1)
                mask = mask.unsqueeze(2)
            emb = emb * mask.to(emb.dtype)
        return self.dropout(emb)


class DebertaV2PreTrainedModel(nn.Module):
    """Weight initialisation and gradient-checkpointing switch shared by the DeBERTa models (reference modeling.py:1042-1068)."""

    base_model_prefix = "deberta"
    _keys_to_ignore_on_load_missing = ["position_ids"]
    _keys_to_ignore_on_load_unexpected = ["position_embeddings"]
    supports_gradient_checkpointing = True
    initializer_range = 0.02

    def _init_weights(self, module):
        if isinstance(module, nn.Linear):
            if not module.weight.is_meta:
                with torch.no_grad():
                    module.weight.normal_(0.0, self.initializer_range)
                    if module.bias is not None:
                        module.bias.zero_()
        elif isinstance(module, nn.Embedding) and not module.weight.is_meta:
            with torch.no_grad():
                module.weight.normal_(0.0, self.initializer_range)
                if module.padding_idx is not None:
                    module.weight[module.padding_idx].zero_()

    def _set_gradient_checkpointing(self, module, value=False):
        if isinstance(module, DebertaV2Encoder):
            module.gradient_checkpointing = value

    def gradient_checkpointing_enable(self, value: bool = True):
        self.apply(lambda m: self._set_gradient_checkpointing(m, value))


class DebertaV2Model(DebertaV2PreTrainedModel):
    """``model(input_ids, attention_mask)`` -> ``BaseModelOutput`` (``.last_hidden_state``); constructor keywords are the fields of the upstream
    ``config.json`` (reference modeling.py:1087-1245).  Unlike the reference default (xxlarge) the conv kernel defaults to the config's value."""

    def __init__(self, _name_or_path="cache/deberta-v-xxlarge", attention_head_size=None, attention_probs_dropout_prob=0.1, conv_act="gelu",
                 conv_kernel_size=3, hidden_act="gelu", hidden_dropout_prob=0.1, hidden_size=1536, initializer_range=0.02, intermediate_size=6144,
                 layer_norm_eps=1e-07, max_position_embeddings=512, max_relative_positions=-1, model_type="deberta-v2", norm_rel_ebd="layer_norm",
                 num_attention_heads=24, num_hidden_layers=48, pad_token_id=0, pooler_dropout=0, pooler_hidden_act="gelu", pooler_hidden_size=1536,
                 pos_att_type=("p2c", "c2p"), position_biased_input=False, position_buckets=256, relative_attention=True, share_att_key=True,
                 type_vocab_size=0, vocab_size=128100, output_attentions=False, output_hidden_states=False, use_return_dict=True, embedding_size=None,
                 dtype=None, device=None, **unused):
        super().__init__()
        if isinstance(pos_att_type, str):
            pos_att_type = [x.strip() for x in pos_att_type.lower().split("|")]
        self.initializer_range = initializer_range
        self.embeddings = DebertaV2Embeddings(max_position_embeddings, position_biased_input, pad_token_id, hidden_size, hidden_dropout_prob,
                                              embedding_size, vocab_size, type_vocab_size, layer_norm_eps, dtype, device)
        self.encoder = DebertaV2Encoder(num_hidden_layers, num_attention_heads, attention_head_size, relative_attention, max_relative_positions,
                                        max_position_embeddings, position_buckets, hidden_size, hidden_act, conv_act, intermediate_size, share_att_key,
                                        list(pos_att_type), norm_rel_ebd, conv_kernel_size, layer_norm_eps, hidden_dropout_prob,
                                        attention_probs_dropout_prob, dtype, device)
        self.z_steps = 0
        self.output_attentions, self.output_hidden_states, self.use_return_dict = output_attentions, output_hidden_states, use_return_dict
        self.hidden_size = hidden_size
        self.apply(self._init_weights)

    def get_input_embeddings(self):
        return self.embeddings.word_embeddings

    def set_input_embeddings(self, new_embeddings):
        self.embeddings.word_embeddings = new_embeddings

    def _prune_heads(self, heads_to_prune):
        raise NotImplementedError("The prune function is not implemented in DeBERTa model.")

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, inputs_embeds=None, output_attentions=None,
                output_hidden_states=None, return_dict=None):
        output_attentions = self.output_attentions if output_attentions is None else output_attentions
        output_hidden_states = self.output_hidden_states if output_hidden_states is None else output_hidden_states
        return_dict = self.use_return_dict if return_dict is None else return_dict
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if input_ids is None and inputs_embeds is None:
            raise ValueError("You have to specify either input_ids or inputs_embeds")
        shape = input_ids.shape if input_ids is not None else inputs_embeds.shape[:-1]
        dev = input_ids.device if input_ids is not None else inputs_embeds.device
        if attention_mask is None:
            attention_mask = torch.ones(tuple(shape), device=dev)
        emb = self.embeddings(input_ids=input_ids, token_type_ids=token_type_ids, position_ids=position_ids, mask=attention_mask,
                              inputs_embeds=inputs_embeds)
        need_layers = output_hidden_states or self.z_steps > 1
        enc = self.encoder(emb, attention_mask, output_hidden_states=need_layers, output_attentions=output_attentions, return_dict=True)
        seq = enc.last_hidden_state
        hidden = enc.hidden_states
        if self.z_steps > 1:              # enhanced mask decoder: re-run the last layer with the previous output as query
            layers = list(hidden)
            kv, query = layers[-2], layers[-1]
            rel, mask4, rel_pos = self.encoder.get_rel_embedding(), self.encoder.get_attention_mask(attention_mask), self.encoder.get_rel_pos(emb)
            for _ in range(self.z_steps - 1):
                query = self.encoder.layer[-1](kv, mask4, output_attentions=False, query_states=query, relative_pos=rel_pos, rel_embeddings=rel)
                layers.append(query)
            seq, hidden = query, tuple(layers)
        hidden = hidden if output_hidden_states else None
        if not return_dict:
            return tuple(v for v in (seq, hidden, enc.attentions) if v is not None)
        return BaseModelOutput(last_hidden_state=seq, hidden_states=hidden, attentions=enc.attentions)


# ----------------------------------------------------------------------------------------------------------------- builders / text encoding
_XXLARGE = dict(attention_head_size=64, attention_probs_dropout_prob=0.1, conv_act="gelu", conv_kernel_size=3, hidden_act="gelu", hidden_dropout_prob=0.1,
                hidden_size=1536, initializer_range=0.02, intermediate_size=6144, layer_norm_eps=1e-07, max_position_embeddings=512,
                max_relative_positions=-1, norm_rel_ebd="layer_norm", num_attention_heads=24, num_hidden_layers=48, pad_token_id=0,
                pos_att_type=["p2c", "c2p"], position_biased_input=False, position_buckets=256, relative_attention=True, share_att_key=True,
                type_vocab_size=0, vocab_size=128100)


def debertav2_xxlarge(**kw):
    return DebertaV2Model(**{**_XXLARGE, **kw})


def debertav2_xlarge(**kw):
    return DebertaV2Model(**{**_XXLARGE, "num_hidden_layers": 24, **kw})


def dict_from_json_file(name):
    with open(os.path.join(name, "config.json"), "r", encoding="utf-8") as reader:
        return json.loads(reader.read())


def get_debertav2_model(name, pretrained=True, dtype=None, device=None, paddle_layout: Optional[bool] = None):
    """Frozen DeBERTa-V2 for the directory ``name``: shape from ``name/config.json`` when present (else xxlarge, the reference's hard-wired
    shape, modeling.py:1248-1290), weights from ``name/debertav2.pd`` when ``pretrained`` (``paddle_layout`` as in ``get_t5_model``).  Eval mode,
    gradients off."""
    if name is None:
        return None
    cfg = dict(_XXLARGE)
    if os.path.isfile(os.path.join(str(name), "config.json")):
        cfg = {**{"conv_kernel_size": 0, "norm_rel_ebd": "none", "share_att_key": False, "relative_attention": False, "pos_att_type": []},
               **dict_from_json_file(name)}
    model = DebertaV2Model(_name_or_path=str(name), dtype=dtype, device=device, **{k: v for k, v in cfg.items() if k != "_name_or_path"})
    if pretrained:
        path = next((p for p in (os.path.join(str(name), f) for f in ("debertav2.pd", "debertav2.pt", "model.pdparams")) if os.path.isfile(p)), None)
        if path is None:
            raise FileNotFoundError(f"no DeBERTa weights under {name} (expected debertav2.pd); pass pretrained=False for random weights")
        ckpt = _read_checkpoint(path)
        sd = ckpt.get("model", ckpt) if isinstance(ckpt, dict) else ckpt
        if paddle_layout is None:
            paddle_layout = any(isinstance(v, np.ndarray) for v in sd.values())
        sd = {k: (torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v) for k, v in sd.items() if hasattr(v, "shape")}
        if paddle_layout:
            linear = {n + ".weight" for n, m in model.named_modules() if isinstance(m, nn.Linear)}
            sd = {k: (v.t().contiguous() if k in linear and v.dim() == 2 else v) for k, v in sd.items()}
        sd = {k: v for k, v in sd.items() if not k.endswith("position_ids")}
        ref = model.embeddings.word_embeddings.weight
        model.load_state_dict({k: v.to(ref.dtype) if v.is_floating_point() else v for k, v in sd.items()}, strict=True)
    model.eval()
    for p in model.parameters():
        p.requires_grad = False
    return model


def debertav2_encode_text(debertav2, texts, tokenizer, return_attn_mask=False):
    """Captions -> frozen DeBERTa features with padded positions zeroed (and the boolean attention mask); reference modeling.py:1300-1313."""
    from ....data.tokenizers import debertav2_tokenize

    token_ids, attn_mask = debertav2_tokenize(texts, tokenizer)
    dev = next(debertav2.parameters()).device
    token_ids, attn_mask = token_ids.to(dev), attn_mask.to(dev)
    debertav2.eval()
    with torch.no_grad():
        encoded = debertav2(input_ids=token_ids, attention_mask=attn_mask).last_hidden_state.detach()
    attn_mask = attn_mask.to(torch.bool)
    encoded = encoded.masked_fill(~attn_mask[:, :, None], 0.0)
    return (encoded, attn_mask) if return_attn_mask else encoded


def get_debertav2_encoded_dim(name):
    if os.path.isfile(os.path.join(str(name), "config.json")):
        return dict_from_json_file(name)["hidden_size"]
    return _XXLARGE["hidden_size"]

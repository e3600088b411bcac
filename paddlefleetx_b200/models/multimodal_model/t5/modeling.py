"""T5 (text tower of Imagen) — reference models/language_model/t5/modeling.py:30-1479.

The module tree and the state-dict keys are the reference's (and the upstream T5 checkpoints'): ``shared``, ``encoder.block.{i}.layer.0.
SelfAttention.{q,k,v,o,relative_attention_bias}``, ``...layer.0.layer_norm``, ``...layer.{-1}.DenseReluDense.{wi | wi_0, wi_1, wo}``,
``encoder.final_layer_norm`` — so converted ``t5.pd`` weights load by name.  A ``T5Stack`` can be an encoder or a decoder (causal
self-attention, cross-attention over encoder states, key / value caches); ``T5EncoderModel`` is what Imagen uses.

B200 notes: RMS layer norm through the fused kernel (``ops.functional.rms_norm``), projections through the framework GEMM, and attention through
the fused SDPA path with the shared relative-position bias as an additive mask whenever the caller does not ask for attention probabilities or
a head mask (T5 scores are un-scaled: ``scale=1``); the explicit softmax path serves inspection, head masking and pruning.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import fields  # noqa: F401  (re-exported: the reference module exposes its own ``fields``)
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ....ops import functional as OF
from ....utils.log import logger
from ...language_model.ernie.model_outputs import BaseModelOutputWithPastAndCrossAttentions, ModelOutput, is_tensor  # noqa: F401


def finfo(dtype):
    """``numpy.finfo`` of a torch floating dtype (the reference clamps fp16 activations at ``finfo(dtype).max - 1000``)."""
    return {torch.float32: np.finfo(np.float32), torch.float16: np.finfo(np.float16), torch.float64: np.finfo(np.float64)}.get(dtype) \
        or torch.finfo(dtype)


# ----------------------------------------------------------------------------------------------------------------- activations
class NewGELUActivation(nn.Module):
    """tanh-approximated GELU (the Google BERT / OpenAI GPT form)."""

    def forward(self, input):
        return F.gelu(input, approximate="tanh")


class GELUActivation(nn.Module):
    """erf GELU; ``use_gelu_python`` spells the formula out instead of calling the fused op."""

    def __init__(self, use_gelu_python: bool = False):
        super().__init__()
        self.act = self._gelu_python if use_gelu_python else F.gelu

    def _gelu_python(self, input):
        return input * 0.5 * (1.0 + torch.erf(input / math.sqrt(2.0)))

    def forward(self, input):
        return self.act(input)


class FastGELUActivation(nn.Module):
    def forward(self, input):
        return 0.5 * input * (1.0 + torch.tanh(input * 0.7978845608 * (1.0 + 0.044715 * input * input)))


class QuickGELUActivation(nn.Module):
    def forward(self, input):
        return input * torch.sigmoid(1.702 * input)


class ClippedGELUActivation(nn.Module):
    """GELU clipped to ``[min, max]`` (keeps the negative lobe representable under quantisation)."""

    def __init__(self, min: float, max: float):
        if min > max:
            raise ValueError(f"min should be < max (got min: {min}, max: {max})")
        super().__init__()
        self.min, self.max = min, max

    def forward(self, x):
        return torch.clip(F.gelu(x), self.min, self.max)


class SiLUActivation(nn.Module):
    def __init__(self):
        super().__init__()
        self.act = F.silu

    def _silu_python(self, input):
        return input * torch.sigmoid(input)

    def forward(self, input):
        return self.act(input)


class MishActivation(nn.Module):
    def __init__(self):
        super().__init__()
        self.act = F.mish

    def _mish_python(self, input):
        return input * torch.tanh(F.softplus(input))

    def forward(self, input):
        return self.act(input)


class LinearActivation(nn.Module):
    def forward(self, input):
        return input


ACT2FN = {
    "gelu": GELUActivation(), "gelu_10": ClippedGELUActivation(-10, 10), "gelu_fast": FastGELUActivation(), "gelu_new": NewGELUActivation(),
    "gelu_python": GELUActivation(use_gelu_python=True), "linear": LinearActivation(), "mish": MishActivation(),
    "quick_gelu": QuickGELUActivation(), "relu": nn.ReLU(), "sigmoid": nn.Sigmoid(), "silu": SiLUActivation(), "swish": SiLUActivation(),
    "tanh": nn.Tanh(),
}


def get_activation(activation_string):
    if activation_string in ACT2FN:
        return ACT2FN[activation_string]
    raise KeyError(f"function {activation_string} not found in ACT2FN mapping {list(ACT2FN.keys())}")


gelu_python, gelu_new, gelu, gelu_fast = (get_activation(n) for n in ("gelu_python", "gelu_new", "gelu", "gelu_fast"))
quick_gelu, silu, mish, linear_act = (get_activation(n) for n in ("quick_gelu", "silu", "mish", "linear"))


# ----------------------------------------------------------------------------------------------------------------- head pruning
def prune_linear_layer(layer: nn.Linear, index: torch.Tensor, dim: int = 0) -> nn.Linear:
    """A new ``nn.Linear`` keeping only ``index`` along ``dim`` of the ``[out, in]`` weight (``dim=0``: output features, bias pruned with
    them; ``dim=1``: input features)."""
    index = index.to(layer.weight.device)
    w = layer.weight.index_select(dim, index).detach().clone()
    out_f, in_f = w.shape
    new = nn.Linear(in_f, out_f, bias=layer.bias is not None, dtype=w.dtype, device=w.device)
    with torch.no_grad():
        new.weight.copy_(w)
        if layer.bias is not None:
            new.bias.copy_(layer.bias.detach() if dim == 1 else layer.bias.detach()[index])
    return new


def find_pruneable_heads_and_indices(heads, n_heads: int, head_size: int, already_pruned_heads):
    """``(heads still to prune, flat indices of the features that stay)``; head numbers refer to the ORIGINAL layout, so each is shifted down
    by the number of already pruned heads in front of it."""
    mask = torch.ones(n_heads, head_size)
    heads = set(heads) - set(already_pruned_heads)
    for head in heads:
        mask[head - sum(1 for h in already_pruned_heads if h < head)] = 0
    index = torch.arange(n_heads * head_size)[mask.view(-1).eq(1)].long()
    return heads, index


# ----------------------------------------------------------------------------------------------------------------- config
class T5Config:
    """Keyword bag with the upstream ``config.json`` fields (reference modeling.py:434-470); unknown keys are kept as attributes too."""

    _DEFAULTS = dict(architectures=None, d_ff=None, d_kv=None, d_model=None, decoder_start_token_id=None, dense_act_fn="gelu_new", eos_token_id=None,
                     feed_forward_proj=None, initializer_factor=None, is_decoder=False, is_encoder_decoder=False, is_gated_act=True,
                     layer_norm_epsilon=None, model_type=None, num_decoder_layers=None, num_heads=None, num_layers=None, output_past=True,
                     pad_token_id=None, relative_attention_max_distance=128, relative_attention_num_buckets=None, tie_word_embeddings=False,
                     transformers_version=None, use_cache=False, vocab_size=None, dropout_rate=None, output_attentions=False,
                     output_hidden_states=False)

    def __init__(self, **kwargs):
        self.use_return_dict = kwargs.pop("return_dict", True)
        for key, default in self._DEFAULTS.items():
            setattr(self, key, kwargs.pop(key, default))
        for key, value in kwargs.items():
            setattr(self, key, value)

    def to_dict(self) -> dict:
        return dict(self.__dict__)

    def model_kwargs(self) -> dict:
        """The constructor arguments of ``T5EncoderModel`` this config determines."""
        keys = ("vocab_size", "d_model", "d_kv", "d_ff", "num_layers", "num_decoder_layers", "num_heads", "relative_attention_num_buckets",
                "dropout_rate", "layer_norm_epsilon", "feed_forward_proj")
        return {k: getattr(self, k) for k in keys if getattr(self, k, None) is not None}


# ----------------------------------------------------------------------------------------------------------------- layers
class T5LayerNorm(nn.Module):
    """RMS norm: scale only, no mean subtraction, no bias; statistics in fp32."""

    def __init__(self, hidden_size, eps=1e-6, dtype=None, device=None):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size, dtype=dtype, device=device))
        self.variance_epsilon = eps

    def forward(self, hidden_states):
        return OF.rms_norm(hidden_states, self.weight, self.variance_epsilon)


def _act_module(name):
    return get_activation(name) if isinstance(name, str) else name


class T5DenseActDense(nn.Module):
    def __init__(self, d_model, d_ff, dropout_rate, dense_act_fn, dtype=None, device=None):
        super().__init__()
        kw = dict(bias=False, dtype=dtype, device=device)
        self.wi, self.wo = nn.Linear(d_model, d_ff, **kw), nn.Linear(d_ff, d_model, **kw)
        self.dropout_p, self.act = dropout_rate, _act_module(dense_act_fn)

    def forward(self, hidden_states):
        h = self.act(OF.linear(hidden_states, self.wi.weight))
        return OF.linear(OF.dropout(h, self.dropout_p, self.training), self.wo.weight)


class T5DenseGatedActDense(nn.Module):
    def __init__(self, d_model, d_ff, dropout_rate, dense_act_fn, dtype=None, device=None):
        super().__init__()
        kw = dict(bias=False, dtype=dtype, device=device)
        self.wi_0, self.wi_1, self.wo = nn.Linear(d_model, d_ff, **kw), nn.Linear(d_model, d_ff, **kw), nn.Linear(d_ff, d_model, **kw)
        self.dropout_p, self.act = dropout_rate, _act_module(dense_act_fn)

    def forward(self, hidden_states):
        h = self.act(OF.linear(hidden_states, self.wi_0.weight)) * OF.linear(hidden_states, self.wi_1.weight)
        return OF.linear(OF.dropout(h, self.dropout_p, self.training), self.wo.weight)


T5DenseReluDense = T5DenseActDense           # earlier name of the un-gated block


class T5LayerFF(nn.Module):
    def __init__(self, d_model, d_ff, dropout_rate, layer_norm_epsilon, feed_forward_proj, dense_act_fn="gelu_new", dtype=None, device=None):
        super().__init__()
        proj = str(feed_forward_proj)
        if proj.startswith("gated-"):                      # "gated-gelu" (T5 v1.1): tanh GELU gate
            self.DenseReluDense = T5DenseGatedActDense(d_model, d_ff, dropout_rate, dense_act_fn if proj == "gated-gelu" else proj[len("gated-"):], dtype,
                                                       device)
        else:
            self.DenseReluDense = T5DenseActDense(d_model, d_ff, dropout_rate, proj, dtype, device)
        self.layer_norm = T5LayerNorm(d_model, layer_norm_epsilon, dtype, device)
        self.dropout_p = dropout_rate

    def forward(self, hidden_states):
        return hidden_states + OF.dropout(self.DenseReluDense(self.layer_norm(hidden_states)), self.dropout_p, self.training)


class T5Attention(nn.Module):
    def __init__(self, is_decoder, relative_attention_num_buckets, d_model, d_kv, num_heads, dropout_rate=0.0, has_relative_attention_bias=False,
                 relative_attention_max_distance=128, dtype=None, device=None):
        super().__init__()
        self.is_decoder, self.has_relative_attention_bias = is_decoder, has_relative_attention_bias
        self.relative_attention_num_buckets, self.relative_attention_max_distance = relative_attention_num_buckets, relative_attention_max_distance
        self.d_model, self.key_value_proj_dim, self.n_heads, self.dropout = d_model, d_kv, num_heads, dropout_rate
        self.inner_dim = num_heads * d_kv
        kw = dict(bias=False, dtype=dtype, device=device)
        # Mesh-TensorFlow initialisation makes scaling before the softmax unnecessary: scores are plain dot products
        self.q, self.k, self.v = nn.Linear(d_model, self.inner_dim, **kw), nn.Linear(d_model, self.inner_dim, **kw), nn.Linear(d_model, self.inner_dim, **kw)
        self.o = nn.Linear(self.inner_dim, d_model, **kw)
        if has_relative_attention_bias:
            self.relative_attention_bias = nn.Embedding(relative_attention_num_buckets, num_heads, dtype=dtype, device=device)
        self.pruned_heads = set()
        self.gradient_checkpointing = False

    def prune_heads(self, heads):
        if not heads:
            return
        heads, index = find_pruneable_heads_and_indices(heads, self.n_heads, self.key_value_proj_dim, self.pruned_heads)
        self.q, self.k, self.v = (prune_linear_layer(m, index) for m in (self.q, self.k, self.v))
        self.o = prune_linear_layer(self.o, index, dim=1)
        # the relative-position bias keeps its original head count: it is computed once by the first block and shared by all blocks, each of
        # which may have pruned different heads — every attention picks the rows of its surviving heads (``_kept_heads``) in ``forward``
        self.n_heads -= len(heads)
        self.inner_dim = self.key_value_proj_dim * self.n_heads
        self.pruned_heads = self.pruned_heads.union(heads)

    @staticmethod
    def _relative_position_bucket(relative_position, bidirectional=True, num_buckets=32, max_distance=128):
        """Signed offset (memory - query) -> bucket id: exact buckets for small offsets, log-spaced up to ``max_distance``, one half of the
        buckets per sign when ``bidirectional`` (encoder), only the past when not (decoder)."""
        buckets = torch.zeros_like(relative_position)
        if bidirectional:
            num_buckets //= 2
            buckets = buckets + (relative_position > 0).long() * num_buckets
            n = relative_position.abs()
        else:
            n = -torch.min(relative_position, torch.zeros_like(relative_position))
        max_exact = num_buckets // 2
        large = max_exact + (torch.log(n.float().clamp(min=1) / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)).long()
        large = torch.min(large, torch.full_like(large, num_buckets - 1))
        return buckets + torch.where(n < max_exact, n, large)

    _bucket = _relative_position_bucket

    def compute_bias(self, query_length, key_length, device=None):
        device = device if device is not None else self.relative_attention_bias.weight.device
        ctx = torch.arange(query_length, dtype=torch.long, device=device)[:, None]
        mem = torch.arange(key_length, dtype=torch.long, device=device)[None, :]
        buckets = self._relative_position_bucket(mem - ctx, bidirectional=not self.is_decoder, num_buckets=self.relative_attention_num_buckets,
                                                 max_distance=self.relative_attention_max_distance)
        return self.relative_attention_bias(buckets).permute(2, 0, 1).unsqueeze(0)          # [1, heads, q, k]

    def forward(self, hidden_states, mask=None, key_value_states=None, position_bias=None, past_key_value=None, layer_head_mask=None,
                query_length=None, use_cache=False, output_attentions=False):
        """Self-attention, or attention over ``key_value_states`` (encoder output).  Returns ``(output, present_key_value or None,
        position_bias[, attention probabilities])``."""
        b, s = hidden_states.shape[:2]
        real_len = s
        if past_key_value is not None:
            assert len(past_key_value) == 2, f"past_key_value should have 2 past states: keys and values. Got {len(past_key_value)} past states"
            real_len += past_key_value[0].shape[2] if query_length is None else query_length
        key_len = real_len if key_value_states is None else key_value_states.shape[1]

        def heads(x):
            return x.view(b, -1, self.n_heads, self.key_value_proj_dim).transpose(1, 2)

        def project(proj, past):
            if key_value_states is None:                       # self-attention: new positions, appended to the cache
                states = heads(OF.linear(hidden_states, proj.weight))
                return states if past is None else torch.cat([past, states], dim=2)
            return heads(OF.linear(key_value_states, proj.weight)) if past is None else past      # cross-attention: computed once

        q = heads(OF.linear(hidden_states, self.q.weight))
        k = project(self.k, None if past_key_value is None else past_key_value[0])
        v = project(self.v, None if past_key_value is None else past_key_value[1])
        if position_bias is None:
            if self.has_relative_attention_bias:
                position_bias = self.compute_bias(real_len, key_len, hidden_states.device).to(q.dtype)
            else:
                position_bias = torch.zeros(1, self.n_heads + len(self.pruned_heads), real_len, key_len, dtype=q.dtype, device=q.device)
            if past_key_value is not None:                     # only the new query rows
                position_bias = position_bias[:, :, -s:, :]
            if mask is not None:
                position_bias = position_bias + mask.to(position_bias.dtype)
        p = self.dropout if self.training else 0.0
        weights = None
        bias = position_bias                                    # shared (original head count); this layer's surviving heads:
        if self.pruned_heads:
            kept = [h for h in range(self.n_heads + len(self.pruned_heads)) if h not in self.pruned_heads]
            bias = position_bias[:, kept]
            if layer_head_mask is not None and layer_head_mask.shape[-3] == len(kept) + len(self.pruned_heads):
                layer_head_mask = layer_head_mask[..., kept, :, :]
        if layer_head_mask is None and not output_attentions:
            out = F.scaled_dot_product_attention(q, k, v, attn_mask=bias.to(q.dtype).expand(b, -1, -1, -1), dropout_p=p, scale=1.0)
        else:
            scores = torch.matmul(q, k.transpose(-1, -2)) + bias
            weights = F.dropout(torch.softmax(scores.float(), dim=-1).to(scores.dtype), p, self.training)
            if layer_head_mask is not None:
                weights = weights * layer_head_mask
            out = torch.matmul(weights, v)
        out = OF.linear(out.transpose(1, 2).reshape(b, -1, self.inner_dim), self.o.weight)
        present = (k, v) if (self.is_decoder and use_cache) else None
        return (out, present, position_bias) + ((weights,) if output_attentions else ())


class T5LayerSelfAttention(nn.Module):
    def __init__(self, is_decoder, relative_attention_num_buckets, d_model, d_kv, num_heads, dropout_rate, layer_norm_epsilon,
                 has_relative_attention_bias=False, dtype=None, device=None):
        super().__init__()
        self.SelfAttention = T5Attention(is_decoder, relative_attention_num_buckets, d_model, d_kv, num_heads, dropout_rate,
                                         has_relative_attention_bias=has_relative_attention_bias, dtype=dtype, device=device)
        self.layer_norm = T5LayerNorm(d_model, layer_norm_epsilon, dtype, device)
        self.dropout_p = dropout_rate

    def forward(self, hidden_states, attention_mask=None, position_bias=None, layer_head_mask=None, past_key_value=None, use_cache=False,
                output_attentions=False):
        att = self.SelfAttention(self.layer_norm(hidden_states), mask=attention_mask, position_bias=position_bias, layer_head_mask=layer_head_mask,
                                 past_key_value=past_key_value, use_cache=use_cache, output_attentions=output_attentions)
        return (hidden_states + OF.dropout(att[0], self.dropout_p, self.training),) + att[1:]


class T5LayerCrossAttention(nn.Module):
    def __init__(self, is_decoder, relative_attention_num_buckets, d_model, d_kv, num_heads, dropout_rate, layer_norm_epsilon, dtype=None, device=None):
        super().__init__()
        self.EncDecAttention = T5Attention(is_decoder, relative_attention_num_buckets, d_model, d_kv, num_heads, dropout_rate,
                                           has_relative_attention_bias=False, dtype=dtype, device=device)
        self.layer_norm = T5LayerNorm(d_model, layer_norm_epsilon, dtype, device)
        self.dropout_p = dropout_rate

    def forward(self, hidden_states, key_value_states, attention_mask=None, position_bias=None, layer_head_mask=None, past_key_value=None,
                use_cache=False, query_length=None, output_attentions=False):
        att = self.EncDecAttention(self.layer_norm(hidden_states), mask=attention_mask, key_value_states=key_value_states, position_bias=position_bias,
                                   layer_head_mask=layer_head_mask, past_key_value=past_key_value, use_cache=use_cache, query_length=query_length,
                                   output_attentions=output_attentions)
        return (hidden_states + OF.dropout(att[0], self.dropout_p, self.training),) + att[1:]


def _clamp_fp16(x):
    """fp16 activations are clamped just inside the representable range (bf16 / fp32 never hit this)."""
    if x.dtype == torch.float16:
        limit = float(torch.finfo(torch.float16).max) - 1000
        x = torch.clamp(x, min=-limit, max=limit)
    return x


class T5Block(nn.Module):
    def __init__(self, is_decoder, relative_attention_num_buckets, feed_forward_proj, d_model, d_kv, num_heads, dropout_rate, layer_norm_epsilon,
                 d_ff, has_relative_attention_bias=False, dtype=None, device=None):
        super().__init__()
        self.is_decoder = is_decoder
        kw = dict(dtype=dtype, device=device)
        self.layer = nn.ModuleList([T5LayerSelfAttention(is_decoder, relative_attention_num_buckets, d_model, d_kv, num_heads, dropout_rate,
                                                         layer_norm_epsilon, has_relative_attention_bias=has_relative_attention_bias, **kw)])
        if is_decoder:
            self.layer.append(T5LayerCrossAttention(is_decoder, relative_attention_num_buckets, d_model, d_kv, num_heads, dropout_rate,
                                                    layer_norm_epsilon, **kw))
        self.layer.append(T5LayerFF(d_model, d_ff, dropout_rate, layer_norm_epsilon, feed_forward_proj, **kw))

    def forward(self, hidden_states, attention_mask=None, position_bias=None, encoder_hidden_states=None, encoder_attention_mask=None,
                encoder_decoder_position_bias=None, layer_head_mask=None, cross_attn_layer_head_mask=None, past_key_value=None, use_cache=False,
                output_attentions=False, return_dict=True):
        """-> ``(hidden, [present_key_values,] self position bias, [self attention,] [cross position bias, [cross attention]])``."""
        self_past = cross_past = None
        if past_key_value is not None:
            if not self.is_decoder:
                logger.warning("`past_key_values` is passed to the encoder. Please make sure this is intended.")
            expected = 2 if encoder_hidden_states is None else 4
            if len(past_key_value) != expected:
                raise ValueError(f"There should be {expected} past states. {'2 (past / key) for cross attention. ' if expected == 4 else ''}"
                                 f"Got {len(past_key_value)} past key / value states")
            self_past, cross_past = past_key_value[:2], (past_key_value[2:] or None)
        sa = self.layer[0](hidden_states, attention_mask=attention_mask, position_bias=position_bias, layer_head_mask=layer_head_mask,
                           past_key_value=self_past, use_cache=use_cache, output_attentions=output_attentions)
        hidden_states, present = _clamp_fp16(sa[0]), sa[1]
        extras = sa[2:]
        if self.is_decoder and encoder_hidden_states is not None:
            qlen = present[0].shape[2] if present is not None else None          # true query length is unknown to the cross attention otherwise
            ca = self.layer[1](hidden_states, key_value_states=encoder_hidden_states, attention_mask=encoder_attention_mask,
                               position_bias=encoder_decoder_position_bias, layer_head_mask=cross_attn_layer_head_mask, past_key_value=cross_past,
                               query_length=qlen, use_cache=use_cache, output_attentions=output_attentions)
            hidden_states = _clamp_fp16(ca[0])
            if present is not None:
                present = present + ca[1]
            extras = extras + ca[2:]
        hidden_states = _clamp_fp16(self.layer[-1](hidden_states))
        return (hidden_states,) + ((present,) if use_cache else ()) + extras


class T5Stack(nn.Module):
    def __init__(self, d_model, num_layers, layer_norm_epsilon, dropout_rate, relative_attention_num_buckets, feed_forward_proj, d_kv, num_heads, d_ff,
                 embed_tokens=None, is_decoder=False, dtype=None, device=None):
        super().__init__()
        # the embedding belongs to whoever built it (``T5EncoderModel.shared``): kept as a plain reference so that it appears once in the
        # state dict; the reference registers it twice and whitelists ``encoder.embed_tokens.weight`` as a missing key
        object.__setattr__(self, "embed_tokens", embed_tokens)
        self.is_decoder, self.num_layers = is_decoder, num_layers
        self.block = nn.ModuleList([T5Block(is_decoder, relative_attention_num_buckets, feed_forward_proj, d_model, d_kv, num_heads, dropout_rate,
                                            layer_norm_epsilon, d_ff, has_relative_attention_bias=(i == 0), dtype=dtype, device=device)
                                    for i in range(num_layers)])
        self.final_layer_norm = T5LayerNorm(d_model, layer_norm_epsilon, dtype, device)
        self.dropout_p = dropout_rate

    def get_input_embeddings(self):
        return self.embed_tokens

    def set_input_embeddings(self, new_embeddings):
        object.__setattr__(self, "embed_tokens", new_embeddings)

    def get_extended_attention_mask(self, attention_mask, input_shape, dtype=torch.float32):
        """``[b, k]`` or ``[b, q, k]`` keep-mask (1 = attend) -> additive ``[b, 1, q | 1, k]`` mask; a decoder also hides the future."""
        if attention_mask.dim() == 3:
            ext = attention_mask[:, None, :, :]
        elif attention_mask.dim() == 2:
            ext = attention_mask[:, None, None, :]
            if self.is_decoder:
                q, k = input_shape[1], attention_mask.shape[1]
                causal = torch.ones(q, k, device=attention_mask.device, dtype=attention_mask.dtype).tril(k - q)
                ext = ext * causal[None, None]
        else:
            raise ValueError(f"Wrong shape for input_ids (shape {tuple(input_shape)}) or attention_mask (shape {tuple(attention_mask.shape)})")
        return (1.0 - ext.to(dtype)) * -10000.0

    def invert_attention_mask(self, encoder_attention_mask, dtype=torch.float32):
        m = encoder_attention_mask
        m = m[:, None, :, :] if m.dim() == 3 else m[:, None, None, :]
        return (1.0 - m.to(dtype)) * -10000.0

    def get_head_mask(self, head_mask, num_hidden_layers, is_attention_chunked=False):
        if head_mask is None:
            return [None] * num_hidden_layers
        head_mask = self._convert_head_mask_to_5d(head_mask, num_hidden_layers)
        return head_mask.unsqueeze(-1) if is_attention_chunked else head_mask

    def _convert_head_mask_to_5d(self, head_mask, num_hidden_layers):
        """``[heads]`` or ``[layers, heads]`` -> ``[layers, batch, heads, q, k]`` (broadcastable)."""
        if head_mask.dim() == 1:
            head_mask = head_mask[None, None, :, None, None].expand(num_hidden_layers, -1, -1, -1, -1)
        elif head_mask.dim() == 2:
            head_mask = head_mask[:, None, :, None, None]
        assert head_mask.dim() == 5, f"head_mask.dim != 5, instead {head_mask.dim()}"
        return head_mask

    def forward(self, input_ids=None, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None, inputs_embeds=None, head_mask=None,
                cross_attn_head_mask=None, past_key_values=None, use_cache=False, output_attentions=False, output_hidden_states=False,
                return_dict=True):
        """Run the stack (encoder, or decoder with cross-attention over ``encoder_hidden_states``) on ids or ready embeddings.  The relative-position
        bias is computed by the first block and shared by the others; ``past_key_values`` / ``use_cache`` extend a decoder incrementally; the
        optional tuples of hidden states / attention probabilities follow the reference's output classes (reference t5/modeling.py:880-1075)."""
        prefix = "decoder_" if self.is_decoder else ""
        if use_cache:
            assert self.is_decoder, f"`use_cache` can only be set to `True` if {type(self).__name__} is used as a decoder"
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError(f"You cannot specify both {prefix}input_ids and {prefix}inputs_embeds at the same time")
        if input_ids is not None:
            input_shape = input_ids.shape
            input_ids = input_ids.reshape(-1, input_shape[-1])
        elif inputs_embeds is not None:
            input_shape = inputs_embeds.shape[:-1]
        else:
            raise ValueError(f"You have to specify either {prefix}input_ids or {prefix}inputs_embeds")
        if inputs_embeds is None:
            assert self.embed_tokens is not None, "You have to initialize the model with valid token embeddings"
            inputs_embeds = self.embed_tokens(input_ids)
        b, s = input_shape
        dev, dtype = inputs_embeds.device, inputs_embeds.dtype
        mask_len = past_key_values[0][0].shape[2] + s if past_key_values is not None else s
        ext_mask = None
        if attention_mask is not None or self.is_decoder:
            if attention_mask is None:
                attention_mask = torch.ones(b, mask_len, device=dev)
            ext_mask = self.get_extended_attention_mask(attention_mask, input_shape, dtype)
        enc_mask = None
        if self.is_decoder and encoder_hidden_states is not None and encoder_attention_mask is not None:
            enc_mask = self.invert_attention_mask(encoder_attention_mask, dtype)
        past_key_values = past_key_values if past_key_values is not None else [None] * len(self.block)
        head_mask = self.get_head_mask(head_mask, self.num_layers)
        cross_attn_head_mask = self.get_head_mask(cross_attn_head_mask, self.num_layers)
        presents = () if use_cache else None
        all_hidden = () if output_hidden_states else None
        all_att = () if output_attentions else None
        all_cross = () if (output_attentions and self.is_decoder) else None
        position_bias = cross_bias = None
        hidden = OF.dropout(inputs_embeds, self.dropout_p, self.training)
        for i, (block, past) in enumerate(zip(self.block, past_key_values)):
            if output_hidden_states:
                all_hidden = all_hidden + (hidden,)
            out = block(hidden, attention_mask=ext_mask, position_bias=position_bias, encoder_hidden_states=encoder_hidden_states,
                        encoder_attention_mask=enc_mask, encoder_decoder_position_bias=cross_bias, layer_head_mask=head_mask[i],
                        cross_attn_layer_head_mask=cross_attn_head_mask[i], past_key_value=past, use_cache=use_cache,
                        output_attentions=output_attentions)
            if not use_cache:
                out = out[:1] + (None,) + out[1:]
            hidden, present = out[:2]
            position_bias = out[2]                               # the first block computes the biases, the others reuse them
            if self.is_decoder and encoder_hidden_states is not None:
                cross_bias = out[4 if output_attentions else 3]
            if use_cache:
                presents = presents + (present,)
            if output_attentions:
                all_att = all_att + (out[3],)
                if self.is_decoder and encoder_hidden_states is not None:
                    all_cross = all_cross + (out[5],)
        hidden = OF.dropout(self.final_layer_norm(hidden), self.dropout_p, self.training)
        if output_hidden_states:
            all_hidden = all_hidden + (hidden,)
        if not return_dict:
            return tuple(v for v in (hidden, presents, all_hidden, all_att, all_cross) if v is not None)
        return BaseModelOutputWithPastAndCrossAttentions(last_hidden_state=hidden, past_key_values=presents, hidden_states=all_hidden,
                                                         attentions=all_att, cross_attentions=all_cross or None)


class T5EncoderModel(nn.Module):
    """Token embedding + encoder stack.  ``model(input_ids, attention_mask)`` returns a ``BaseModelOutputWithPastAndCrossAttentions``
    (``.last_hidden_state``), or a tuple with ``return_dict=False`` (reference modeling.py:1318-1409)."""

    authorized_missing_keys = [r"encoder.embed_tokens.weight"]

    def __init__(self, vocab_size=32128, d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_decoder_layers=12, num_heads=12,
                 relative_attention_num_buckets=32, dropout_rate=0.1, layer_norm_epsilon=1e-06, feed_forward_proj="relu", pad_token_id=0,
                 dtype=None, device=None, **unused):
        super().__init__()
        self.config = dict(vocab_size=vocab_size, d_model=d_model, d_kv=d_kv, d_ff=d_ff, num_layers=num_layers, num_heads=num_heads,
                           relative_attention_num_buckets=relative_attention_num_buckets, dropout_rate=dropout_rate,
                           layer_norm_epsilon=layer_norm_epsilon, feed_forward_proj=feed_forward_proj)
        self.shared = nn.Embedding(vocab_size, d_model, dtype=dtype, device=device)
        self.encoder = T5Stack(d_model, num_layers, layer_norm_epsilon, dropout_rate, relative_attention_num_buckets, feed_forward_proj, d_kv, num_heads,
                               d_ff, embed_tokens=self.shared, is_decoder=False, dtype=dtype, device=device)
        self.pad_token_id, self.d_model = pad_token_id, d_model
        self._register_load_state_dict_pre_hook(self._drop_duplicate_embedding)

    def get_input_embeddings(self):
        return self.shared

    def set_input_embeddings(self, new_embeddings):
        self.shared = new_embeddings
        self.encoder.set_input_embeddings(new_embeddings)

    def get_encoder(self):
        return self.encoder

    def _prune_heads(self, heads_to_prune):
        """``{layer index: [heads]}``."""
        for layer, heads in heads_to_prune.items():
            self.encoder.block[layer].layer[0].SelfAttention.prune_heads(heads)

    @staticmethod
    def _drop_duplicate_embedding(state_dict, prefix, *_):
        state_dict.pop(prefix + "encoder.embed_tokens.weight", None)      # upstream checkpoints carry ``shared.weight`` a second time under this name

    def forward(self, input_ids=None, attention_mask=None, head_mask=None, inputs_embeds=None, output_attentions=None, output_hidden_states=None,
                return_dict=None):
        return self.encoder(input_ids=input_ids, attention_mask=attention_mask, inputs_embeds=inputs_embeds, head_mask=head_mask,
                            output_attentions=bool(output_attentions), output_hidden_states=bool(output_hidden_states),
                            return_dict=True if return_dict is None else return_dict)


# ----------------------------------------------------------------------------------------------------------------- builders / text encoding
_T5 = {
    "t5-small": dict(d_model=512, d_kv=64, d_ff=2048, num_layers=6, num_heads=8),
    "t5-base": dict(d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_heads=12),
    "t5-large": dict(d_model=1024, d_kv=64, d_ff=4096, num_layers=24, num_heads=16),
    "t5-3b": dict(d_model=1024, d_kv=128, d_ff=16384, num_layers=24, num_heads=32),
    "t5-11b": dict(d_model=1024, d_kv=128, d_ff=65536, num_layers=24, num_heads=128),
}


def T5Model(config, **kw):
    """Encoder built from a ``config.json``-style dict (or a ``T5Config``)."""
    config = config if isinstance(config, T5Config) else T5Config(**dict(config))
    return T5EncoderModel(**{**config.model_kwargs(), **kw})


def dict_from_json_file(name):
    with open(os.path.join(name, "config.json"), "r", encoding="utf-8") as reader:
        return json.loads(reader.read())


def _preset_of(name: str) -> Optional[dict]:
    base = os.path.basename(str(name).rstrip("/")).lower()
    return next((dict(v) for k, v in _T5.items() if base == k or base.endswith(k)), None)


def _read_checkpoint(path):
    """``torch.save`` archives, or the plain pickles of numpy arrays that ``paddle.save`` writes."""
    try:
        return torch.load(path, map_location="cpu", weights_only=False)
    except Exception:                       # noqa: BLE001 - not a torch archive: try the pickle
        import pickle

        with open(path, "rb") as f:
            return pickle.load(f, encoding="latin1")


def get_t5_model(name, pretrained=True, dtype=None, device=None, paddle_layout: Optional[bool] = None):
    """Frozen T5 encoder for the directory ``name`` (e.g. ``t5/t5-11b``): shape from ``name/config.json`` when present, else from the preset
    the directory is named after; weights from ``name/t5.pd`` (``{"model": state_dict}`` or a bare state dict) when ``pretrained``.
    ``paddle_layout``: Linear weights stored ``[in, out]`` (Paddle) instead of ``[out, in]``; by default assumed exactly when the file holds
    numpy arrays, i.e. was written by ``paddle.save``.  Returned in eval mode with gradients off (reference modeling.py:1417-1440)."""
    if os.path.isfile(os.path.join(str(name), "config.json")):
        model = T5Model(dict_from_json_file(name), dropout_rate=0.0, dtype=dtype, device=device)
    else:
        shape = _preset_of(name)
        if shape is None:
            raise FileNotFoundError(f"{name}/config.json not found and {name!r} names no T5 preset ({sorted(_T5)})")
        model = T5EncoderModel(vocab_size=32128, relative_attention_num_buckets=32, dropout_rate=0.0, layer_norm_epsilon=1e-6, feed_forward_proj="relu",
                               dtype=dtype, device=device, **shape)
    if pretrained:
        path = next((p for p in (os.path.join(str(name), f) for f in ("t5.pd", "t5.pt", "model.pdparams")) if os.path.isfile(p)), None)
        if path is None:
            raise FileNotFoundError(f"no T5 weights under {name} (expected t5.pd); pass pretrained=False for random weights")
        ckpt = _read_checkpoint(path)
        sd = ckpt.get("model", ckpt) if isinstance(ckpt, dict) else ckpt
        if paddle_layout is None:
            paddle_layout = any(isinstance(v, np.ndarray) for v in sd.values())
        sd = {k: (torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v) for k, v in sd.items() if hasattr(v, "shape")}
        if paddle_layout:
            linear = tuple(f".{n}.weight" for n in ("q", "k", "v", "o", "wi", "wi_0", "wi_1", "wo"))
            sd = {k: (v.t().contiguous() if k.endswith(linear) and v.dim() == 2 else v) for k, v in sd.items()}
        ref = model.shared.weight
        model.load_state_dict({k: v.to(ref.dtype) for k, v in sd.items()}, strict=True)
    model.eval()
    for p in model.parameters():
        p.requires_grad = False
    return model


def t5_encoder(name: str = "t5-11b", **kw) -> T5EncoderModel:
    return T5EncoderModel(**{**_T5[name], **kw})


def t5_11b(**kw):
    """The Imagen text tower: 24 layers, 128 heads of 128, 65536-wide ReLU FFN, dropout off (reference modeling.py:1442-1454)."""
    return t5_encoder("t5-11b", **{**dict(vocab_size=32128, relative_attention_num_buckets=32, dropout_rate=0.0, layer_norm_epsilon=1e-6,
                                          feed_forward_proj="relu"), **kw})


def t5_3b(**kw): return t5_encoder("t5-3b", **kw)
def t5_large(**kw): return t5_encoder("t5-large", **kw)
def t5_base(**kw): return t5_encoder("t5-base", **kw)
def t5_small(**kw): return t5_encoder("t5-small", **kw)


def t5_encode_text(t5, texts, tokenizer, return_attn_mask=False):
    """Captions -> frozen T5 features ``[b, s, d_model]`` (and the attention mask): tokenises with ``t5_tokenize``, runs the encoder in eval mode
    without gradients (reference modeling.py:1464-1475)."""
    from ....data.tokenizers import t5_tokenize

    token_ids, attn_mask = t5_tokenize(texts, tokenizer)
    dev = next(t5.parameters()).device
    token_ids, attn_mask = token_ids.to(dev), attn_mask.to(dev)
    t5.eval()
    with torch.no_grad():
        text_features = t5(input_ids=token_ids, attention_mask=attn_mask).last_hidden_state.detach()
    return (text_features, attn_mask) if return_attn_mask else text_features


def get_encoded_dim(name):
    """Feature width of the T5 stored under ``name`` (``config.json``), or of the preset the directory is named after."""
    if os.path.isfile(os.path.join(str(name), "config.json")):
        return dict_from_json_file(name)["d_model"]
    shape = _preset_of(name)
    if shape is None:
        raise FileNotFoundError(f"{name}/config.json not found and {name!r} names no T5 preset")
    return shape["d_model"]

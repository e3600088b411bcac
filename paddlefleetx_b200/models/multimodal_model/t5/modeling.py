"""T5 encoder (text tower of Imagen) — reference models/language_model/t5/modeling.py:434-1479: RMS layer norm without
bias, relative-position-bucket attention bias shared from layer 0, (gated-)GELU / ReLU feed-forward, ``T5EncoderModel``
and the ``t5_11b`` / ``t5_*`` presets.  Linear layers run through the framework GEMM; norms through the fused RMSNorm."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ....ops import functional as OF


class T5LayerNorm(nn.Module):
    def __init__(self, hidden, eps=1e-6, dtype=None, device=None):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden, dtype=dtype, device=device))
        self.eps = eps

    def forward(self, x):
        return OF.rms_norm(x, self.weight, self.eps)


class T5DenseReluDense(nn.Module):
    def __init__(self, d_model, d_ff, dropout, gated: bool, act: str, dtype=None, device=None):
        super().__init__()
        kw = dict(bias=False, dtype=dtype, device=device)
        self.gated, self.act, self.dropout = gated, act, dropout
        if gated:
            self.wi_0 = nn.Linear(d_model, d_ff, **kw); self.wi_1 = nn.Linear(d_model, d_ff, **kw)
        else:
            self.wi = nn.Linear(d_model, d_ff, **kw)
        self.wo = nn.Linear(d_ff, d_model, **kw)

    def _act(self, x):
        return F.gelu(x, approximate="tanh") if "gelu" in self.act else F.relu(x)

    def forward(self, x):
        if self.gated:
            h = self._act(OF.linear(x, self.wi_0.weight)) * OF.linear(x, self.wi_1.weight)
        else:
            h = self._act(OF.linear(x, self.wi.weight))
        return OF.linear(OF.dropout(h, self.dropout, self.training), self.wo.weight)


class T5Attention(nn.Module):
    def __init__(self, d_model, d_kv, num_heads, dropout, has_relative_attention_bias, num_buckets=32, max_distance=128, dtype=None, device=None):
        super().__init__()
        self.h, self.d_kv, self.inner = num_heads, d_kv, num_heads * d_kv
        kw = dict(bias=False, dtype=dtype, device=device)
        self.q, self.k, self.v = nn.Linear(d_model, self.inner, **kw), nn.Linear(d_model, self.inner, **kw), nn.Linear(d_model, self.inner, **kw)
        self.o = nn.Linear(self.inner, d_model, **kw)
        self.dropout, self.num_buckets, self.max_distance = dropout, num_buckets, max_distance
        self.has_bias = has_relative_attention_bias
        if has_relative_attention_bias:
            self.relative_attention_bias = nn.Embedding(num_buckets, num_heads, dtype=dtype, device=device)

    @staticmethod
    def _bucket(rel, num_buckets=32, max_distance=128):
        num_buckets //= 2
        ret = (rel > 0).long() * num_buckets
        n = rel.abs()
        max_exact = num_buckets // 2
        large = max_exact + (torch.log(n.float().clamp(min=1) / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)).long()
        large = large.clamp(max=num_buckets - 1)
        return ret + torch.where(n < max_exact, n, large)

    def compute_bias(self, q_len, k_len, device):
        ctx = torch.arange(q_len, device=device)[:, None]
        mem = torch.arange(k_len, device=device)[None, :]
        buckets = self._bucket(mem - ctx, self.num_buckets, self.max_distance)
        return self.relative_attention_bias(buckets).permute(2, 0, 1).unsqueeze(0)     # [1, h, q, k]

    def forward(self, x, mask=None, position_bias=None):
        b, s, _ = x.shape
        q, k, v = (OF.linear(x, w.weight).view(b, s, self.h, self.d_kv) for w in (self.q, self.k, self.v))
        if position_bias is None:
            position_bias = self.compute_bias(s, s, x.device) if self.has_bias else torch.zeros(1, self.h, s, s, device=x.device, dtype=x.dtype)
            if mask is not None:
                position_bias = position_bias + mask
        # T5 uses un-scaled dot products (scale folded into the init)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=position_bias.to(q.dtype),
                                           dropout_p=self.dropout if self.training else 0.0, scale=1.0)
        return OF.linear(o.transpose(1, 2).reshape(b, s, self.inner), self.o.weight), position_bias


class T5Block(nn.Module):
    def __init__(self, cfg, has_relative_attention_bias, dtype=None, device=None):
        super().__init__()
        self.ln1 = T5LayerNorm(cfg["d_model"], cfg["layer_norm_epsilon"], dtype, device)
        self.attn = T5Attention(cfg["d_model"], cfg["d_kv"], cfg["num_heads"], cfg["dropout_rate"], has_relative_attention_bias,
                                cfg["relative_attention_num_buckets"], cfg.get("relative_attention_max_distance", 128), dtype, device)
        self.ln2 = T5LayerNorm(cfg["d_model"], cfg["layer_norm_epsilon"], dtype, device)
        self.ff = T5DenseReluDense(cfg["d_model"], cfg["d_ff"], cfg["dropout_rate"], "gated" in cfg["feed_forward_proj"], cfg["feed_forward_proj"], dtype, device)
        self.dropout = cfg["dropout_rate"]

    def forward(self, x, mask=None, position_bias=None):
        a, position_bias = self.attn(self.ln1(x), mask, position_bias)
        x = x + OF.dropout(a, self.dropout, self.training)
        x = x + OF.dropout(self.ff(self.ln2(x)), self.dropout, self.training)
        return x, position_bias


_T5 = {
    "t5-small": dict(d_model=512, d_kv=64, d_ff=2048, num_layers=6, num_heads=8),
    "t5-base": dict(d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_heads=12),
    "t5-large": dict(d_model=1024, d_kv=64, d_ff=4096, num_layers=24, num_heads=16),
    "t5-3b": dict(d_model=1024, d_kv=128, d_ff=16384, num_layers=24, num_heads=32),
    "t5-11b": dict(d_model=1024, d_kv=128, d_ff=65536, num_layers=24, num_heads=128),
}


class T5EncoderModel(nn.Module):
    def __init__(self, vocab_size=32128, d_model=512, d_kv=64, d_ff=2048, num_layers=6, num_heads=8, relative_attention_num_buckets=32,
                 dropout_rate=0.1, layer_norm_epsilon=1e-6, feed_forward_proj="relu", pad_token_id=0, dtype=None, device=None, **unused):
        super().__init__()
        cfg = dict(d_model=d_model, d_kv=d_kv, d_ff=d_ff, num_heads=num_heads, relative_attention_num_buckets=relative_attention_num_buckets,
                   dropout_rate=dropout_rate, layer_norm_epsilon=layer_norm_epsilon, feed_forward_proj=feed_forward_proj)
        self.config = dict(cfg, vocab_size=vocab_size, num_layers=num_layers)
        self.shared = nn.Embedding(vocab_size, d_model, dtype=dtype, device=device)
        self.block = nn.ModuleList([T5Block(cfg, i == 0, dtype, device) for i in range(num_layers)])
        self.final_layer_norm = T5LayerNorm(d_model, layer_norm_epsilon, dtype, device)
        self.dropout, self.pad_token_id, self.d_model = dropout_rate, pad_token_id, d_model

    def forward(self, input_ids, attention_mask=None):
        x = OF.dropout(self.shared(input_ids), self.dropout, self.training)
        mask = None
        if attention_mask is not None:
            mask = (1.0 - attention_mask[:, None, None, :].to(x.dtype)) * -1e4
        bias = None
        for blk in self.block:
            x, bias = blk(x, mask, bias)
        return OF.dropout(self.final_layer_norm(x), self.dropout, self.training)


def t5_encoder(name: str = "t5-11b", **kw) -> T5EncoderModel:
    return T5EncoderModel(**{**_T5[name], **kw})


def t5_11b(**kw): return t5_encoder("t5-11b", **kw)
def t5_3b(**kw): return t5_encoder("t5-3b", **kw)
def t5_large(**kw): return t5_encoder("t5-large", **kw)
def t5_base(**kw): return t5_encoder("t5-base", **kw)
def t5_small(**kw): return t5_encoder("t5-small", **kw)

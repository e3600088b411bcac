"""DeBERTa-V2 encoder (alternative Imagen text tower) — reference models/language_model/debertav2/modeling.py:57-1323:
disentangled attention (content-to-position and position-to-content terms over log-bucketed relative positions), shared
relative embeddings with layer norm, optional convolution after the first layer, ``DebertaV2Model``."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ....ops import functional as OF


def make_log_bucket_position(rel, bucket_size, max_position):
    sign = torch.sign(rel)
    mid = bucket_size // 2
    abs_pos = torch.where((rel < mid) & (rel > -mid), torch.full_like(rel, mid - 1), rel.abs())
    log_pos = torch.ceil(torch.log(abs_pos.float() / mid) / math.log((max_position - 1) / mid) * (mid - 1)) + mid
    return torch.where(abs_pos <= mid, rel.float(), log_pos * sign).long()


def build_relative_position(q_len, k_len, bucket_size=-1, max_position=-1, device=None):
    rel = torch.arange(q_len, device=device)[:, None] - torch.arange(k_len, device=device)[None, :]
    if bucket_size > 0 and max_position > 0:
        rel = make_log_bucket_position(rel, bucket_size, max_position)
    return rel.unsqueeze(0)


class DisentangledSelfAttention(nn.Module):
    def __init__(self, cfg, dtype=None, device=None):
        super().__init__()
        h, heads = cfg["hidden_size"], cfg["num_attention_heads"]
        self.heads, self.hd = heads, h // heads
        kw = dict(dtype=dtype, device=device)
        self.query_proj, self.key_proj, self.value_proj = nn.Linear(h, h, **kw), nn.Linear(h, h, **kw), nn.Linear(h, h, **kw)
        self.pos_att_type = cfg.get("pos_att_type", ["p2c", "c2p"])
        self.position_buckets = cfg.get("position_buckets", 256)
        self.max_relative_positions = cfg.get("max_relative_positions", -1)
        if self.max_relative_positions < 1:
            self.max_relative_positions = cfg.get("max_position_embeddings", 512)
        self.pos_ebd_size = self.position_buckets if self.position_buckets > 0 else self.max_relative_positions
        self.dropout = cfg.get("attention_probs_dropout_prob", 0.1)

    def _heads(self, x):
        b, s, _ = x.shape
        return x.view(b, s, self.heads, self.hd).transpose(1, 2)           # [b, h, s, d]

    def forward(self, x, mask, rel_embeddings, relative_pos=None):
        b, s, _ = x.shape
        q, k, v = self._heads(self.query_proj(x)), self._heads(self.key_proj(x)), self._heads(self.value_proj(x))
        scale_factor = 1 + ("c2p" in self.pos_att_type) + ("p2c" in self.pos_att_type)
        scale = 1.0 / math.sqrt(self.hd * scale_factor)
        scores = torch.matmul(q, k.transpose(-1, -2)) * scale
        if relative_pos is None:
            relative_pos = build_relative_position(s, s, self.position_buckets, self.max_relative_positions, x.device)
        span = self.pos_ebd_size
        rel = rel_embeddings[: 2 * span].unsqueeze(0)
        pos_k = self._heads(self.key_proj(rel))                                   # [1, h, 2span, d]
        pos_q = self._heads(self.query_proj(rel))
        idx = (relative_pos + span).clamp(0, 2 * span - 1)                         # [1, s, s]
        if "c2p" in self.pos_att_type:
            c2p = torch.matmul(q, pos_k.transpose(-1, -2))                         # [b, h, s, 2span]
            scores = scores + torch.gather(c2p, -1, idx.unsqueeze(1).expand(b, self.heads, s, s)) * scale
        if "p2c" in self.pos_att_type:
            p2c = torch.matmul(k, pos_q.transpose(-1, -2))                         # [b, h, s, 2span]
            idx_t = (-relative_pos + span).clamp(0, 2 * span - 1)
            scores = scores + torch.gather(p2c, -1, idx_t.unsqueeze(1).expand(b, self.heads, s, s)).transpose(-1, -2) * scale
        if mask is not None:
            scores = scores + mask
        probs = OF.dropout(torch.softmax(scores.float(), -1).to(x.dtype), self.dropout, self.training)
        return torch.matmul(probs, v).transpose(1, 2).reshape(b, s, -1)


class DebertaV2Layer(nn.Module):
    def __init__(self, cfg, dtype=None, device=None):
        super().__init__()
        h, kw = cfg["hidden_size"], dict(dtype=dtype, device=device)
        self.attn = DisentangledSelfAttention(cfg, dtype, device)
        self.attn_out, self.attn_ln = nn.Linear(h, h, **kw), nn.LayerNorm(h, cfg.get("layer_norm_eps", 1e-7), **kw)
        self.inter, self.out = nn.Linear(h, cfg["intermediate_size"], **kw), nn.Linear(cfg["intermediate_size"], h, **kw)
        self.out_ln = nn.LayerNorm(h, cfg.get("layer_norm_eps", 1e-7), **kw)
        self.dropout = cfg.get("hidden_dropout_prob", 0.1)

    def forward(self, x, mask, rel_embeddings, relative_pos=None):
        a = self.attn(x, mask, rel_embeddings, relative_pos)
        x = self.attn_ln(x + OF.dropout(self.attn_out(a), self.dropout, self.training))
        f = self.out(F.gelu(self.inter(x)))
        return self.out_ln(x + OF.dropout(f, self.dropout, self.training))


class ConvLayer(nn.Module):
    def __init__(self, cfg, dtype=None, device=None):
        super().__init__()
        k = cfg.get("conv_kernel_size", 3)
        h = cfg["hidden_size"]
        self.conv = nn.Conv1d(h, h, k, padding=(k - 1) // 2, groups=cfg.get("conv_groups", 1), dtype=dtype, device=device)
        self.ln = nn.LayerNorm(h, cfg.get("layer_norm_eps", 1e-7), dtype=dtype, device=device)
        self.act, self.dropout = cfg.get("conv_act", "tanh"), cfg.get("hidden_dropout_prob", 0.1)

    def forward(self, hidden, residual, input_mask):
        out = self.conv(hidden.transpose(1, 2)).transpose(1, 2)
        out = out.masked_fill(~input_mask.bool().unsqueeze(-1), 0)
        out = torch.tanh(out) if self.act == "tanh" else F.gelu(out)
        out = self.ln(residual + OF.dropout(out, self.dropout, self.training))
        return out * input_mask.unsqueeze(-1).to(out.dtype)


class DebertaV2Model(nn.Module):
    def __init__(self, vocab_size=128100, hidden_size=1536, num_hidden_layers=24, num_attention_heads=24, intermediate_size=6144,
                 max_position_embeddings=512, position_buckets=256, relative_attention=True, norm_rel_ebd="layer_norm", conv_kernel_size=0,
                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, layer_norm_eps=1e-7, pad_token_id=0, pos_att_type=("p2c", "c2p"),
                 position_biased_input=False, type_vocab_size=0, dtype=None, device=None, **unused):
        super().__init__()
        cfg = dict(hidden_size=hidden_size, num_attention_heads=num_attention_heads, intermediate_size=intermediate_size,
                   max_position_embeddings=max_position_embeddings, position_buckets=position_buckets, max_relative_positions=-1,
                   hidden_dropout_prob=hidden_dropout_prob, attention_probs_dropout_prob=attention_probs_dropout_prob, layer_norm_eps=layer_norm_eps,
                   pos_att_type=list(pos_att_type), conv_kernel_size=conv_kernel_size)
        kw = dict(dtype=dtype, device=device)
        self.word_embeddings = nn.Embedding(vocab_size, hidden_size, padding_idx=pad_token_id, **kw)
        self.position_biased_input = position_biased_input
        if position_biased_input:
            self.position_embeddings = nn.Embedding(max_position_embeddings, hidden_size, **kw)
        self.embed_ln = nn.LayerNorm(hidden_size, layer_norm_eps, **kw)
        self.layer = nn.ModuleList([DebertaV2Layer(cfg, dtype, device) for _ in range(num_hidden_layers)])
        span = position_buckets if position_buckets > 0 else max_position_embeddings
        self.rel_embeddings = nn.Embedding(span * 2, hidden_size, **kw)
        self.rel_ln = nn.LayerNorm(hidden_size, layer_norm_eps, **kw) if norm_rel_ebd == "layer_norm" else None
        self.conv = ConvLayer(cfg, dtype, device) if conv_kernel_size > 0 else None
        self.dropout, self.hidden_size = hidden_dropout_prob, hidden_size

    def forward(self, input_ids, attention_mask=None):
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        x = self.word_embeddings(input_ids)
        if self.position_biased_input:
            x = x + self.position_embeddings(torch.arange(input_ids.shape[1], device=input_ids.device))[None]
        x = OF.dropout(self.embed_ln(x) * attention_mask.unsqueeze(-1).to(x.dtype), self.dropout, self.training)
        ext = attention_mask[:, None, None, :] * attention_mask[:, None, :, None]
        mask = (1.0 - ext.to(x.dtype)) * -1e4
        rel = self.rel_embeddings.weight
        if self.rel_ln is not None:
            rel = self.rel_ln(rel)
        h = x
        for i, layer in enumerate(self.layer):
            out = layer(h, mask, rel)
            if i == 0 and self.conv is not None:
                out = self.conv(x, out, attention_mask)
            h = out
        return h


def debertav2_xxlarge(**kw): return DebertaV2Model(**{**dict(hidden_size=1536, num_hidden_layers=48, num_attention_heads=24, intermediate_size=6144, conv_kernel_size=3), **kw})
def debertav2_xlarge(**kw): return DebertaV2Model(**{**dict(hidden_size=1536, num_hidden_layers=24, num_attention_heads=24, intermediate_size=6144, conv_kernel_size=3), **kw})

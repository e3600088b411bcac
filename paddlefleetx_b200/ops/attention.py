"""Attention front-end.

``attention(q, k, v)`` takes ``[b, s, heads, d]`` tensors.  Path selection:
  * our tcgen05 / TMEM flash forward (``csrc/attention_fwd_sm100.cu``) on the no-grad paths (generation prefill, evaluation,
    inference of the vision / text encoders) when no dropout and no explicit mask are requested and head_dim is 64 or 128;
    single-query decode steps over the static KV cache use ``csrc/attention_decode.cu`` (see models/.../gpt/model.py),
  * otherwise the library SDPA (cuDNN / FlashAttention-2 inside PyTorch) — a *library* call, reported as such
    by the launch accounting (it does not count towards ``gpu_launches``),
  * the unfused reference path (QK^T -> fused causal softmax -> PV) used when an explicit mask is given or
    ``use_flash_attn=False`` — this mirrors reference ``core_attn`` (hybrid_model.py:303-346).
"""
from __future__ import annotations

from typing import Optional

import os

import torch
import torch.nn.functional as F

# 0 = never, 1 = where it is at least as fast as the library kernel (short sequences: prefill / few-hundred-token inputs), 2 = always.
# Measured on B200 (profiles/README.md): 373 TFLOP/s at S=1024 causal D=128 vs 839 for cuDNN — correct, but its per-tile softmax path is
# still ~3x the MMA time, so long sequences stay on the library until the ping-pong version lands.
_NATIVE_FWD = int(os.environ.get("PFX_NATIVE_ATTN_FWD", "1"))
_NATIVE_FWD_MAX_SEQ = 256


def _native_fwd_ok(q, k, v, dropout_p, attn_mask, causal) -> bool:
    if not _NATIVE_FWD or attn_mask is not None or dropout_p != 0.0 or not q.is_cuda or q.dtype != torch.bfloat16:
        return False
    if _NATIVE_FWD == 1 and max(q.shape[1], k.shape[1]) > _NATIVE_FWD_MAX_SEQ:
        return False
    if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad):
        return False                      # training keeps the library forward/backward pair
    if q.shape[-1] not in (64, 128) or q.shape[1] < 16 or (causal and k.shape[1] < q.shape[1]):
        return False
    from . import _native

    return _native.available()


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = True, dropout_p: float = 0.0,
              scale: Optional[float] = None, attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    if _native_fwd_ok(q, k, v, dropout_p, attn_mask, causal):
        from . import _native
        from . import functional as OF

        OF._count()
        out, _ = _native.require().attention_fwd(q.contiguous(), k.contiguous(), v.contiguous(), bool(causal),
                                                 float(scale if scale is not None else q.shape[-1] ** -0.5))
        return out
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

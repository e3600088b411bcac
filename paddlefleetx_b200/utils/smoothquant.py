"""Shift-SmoothQuant post-training quantisation for export (new capability: README-only in the reference).

SmoothQuant (Xiao et al. 2022) migrates activation outliers into the weights with a per-input-channel factor
``s_j = max|X_j|^alpha / max|W_j|^(1-alpha)`` so that ``(X / s)(W * s)^T == X W^T`` while both factors quantise well to int8.
The *shift* variant first centres every activation channel, ``X' = X - z`` with ``z_j = (max_j + min_j) / 2``, folding
``z W^T`` into the layer bias — asymmetric channels (post-GELU, post-LayerNorm with bias) then use the full int8 range.

``calibrate`` collects per-channel min / max over calibration batches with forward hooks; ``smooth_and_quantize`` rewrites
every ``nn.Linear``-like layer into ``ShiftSmoothInt8Linear`` running the int8 tcgen05 GEMM.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import torch
import torch.nn as nn

from ..ops.quant import Int8Linear, quantize_weight_int8

_TYPES = ("Linear", "ColumnParallelLinear", "RowParallelLinear")


class ActStats:
    def __init__(self):
        self.min: Optional[torch.Tensor] = None
        self.max: Optional[torch.Tensor] = None

    def update(self, x: torch.Tensor) -> None:
        x2 = x.detach().reshape(-1, x.shape[-1]).float()
        mn, mx = x2.amin(0), x2.amax(0)
        self.min = mn if self.min is None else torch.minimum(self.min, mn)
        self.max = mx if self.max is None else torch.maximum(self.max, mx)


def calibrate(model: nn.Module, batches: Iterable, forward_fn=None) -> Dict[str, ActStats]:
    stats: Dict[str, ActStats] = {}
    hooks = []
    for name, mod in model.named_modules():
        if mod.__class__.__name__ in _TYPES and getattr(mod, "world", 1) == 1:
            stats[name] = ActStats()
            hooks.append(mod.register_forward_pre_hook(lambda m, a, n=name: stats[n].update(a[0])))
    model.eval()
    with torch.no_grad():
        for b in batches:
            forward_fn(model, b) if forward_fn is not None else model(*b) if isinstance(b, (tuple, list)) else model(b)
    for h in hooks:
        h.remove()
    return stats


class ShiftSmoothInt8Linear(nn.Module):
    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor], stats: ActStats, alpha: float = 0.5, shift: bool = True):
        super().__init__()
        w = weight.detach().float()
        z = ((stats.max + stats.min) / 2).to(w.device) if shift else torch.zeros(w.shape[1], device=w.device)
        amax = torch.maximum((stats.max.to(w.device) - z).abs(), (stats.min.to(w.device) - z).abs()).clamp(min=1e-5)
        wmax = w.abs().amax(0).clamp(min=1e-5)
        s = (amax.pow(alpha) / wmax.pow(1 - alpha)).clamp(min=1e-5)
        b = (bias.detach().float() if bias is not None else torch.zeros(w.shape[0], device=w.device)) + w @ z     # fold the shift
        self.register_buffer("shift", z.to(weight.dtype))
        self.inner = Int8Linear.from_float(w, b.to(torch.bfloat16), smooth=s)

    def forward(self, x):
        return self.inner(x - self.shift.to(x.dtype))


def smooth_and_quantize(model: nn.Module, stats: Dict[str, ActStats], alpha: float = 0.5, shift: bool = True, skip=("score", "head")) -> nn.Module:
    for name, mod in list(model.named_modules()):
        for child_name, child in list(mod.named_children()):
            full = f"{name}.{child_name}" if name else child_name
            if full in stats and stats[full].max is not None and isinstance(child, nn.Linear) and not any(s in child_name for s in skip):
                setattr(mod, child_name, ShiftSmoothInt8Linear(child.weight.data, child.bias, stats[full], alpha, shift))
    return model

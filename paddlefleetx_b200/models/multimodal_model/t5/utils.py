"""In-place initialisers used around the T5 encoder (reference ppfleetx/models/language_model/t5/utils.py:24-50): fill a tensor, or a
layer's ``weight`` / ``bias``, with a constant or with normal noise.  All of them write under ``no_grad`` and return their argument."""
from __future__ import annotations

import torch


@torch.no_grad()
def constant_(x: torch.Tensor, value) -> torch.Tensor:
    return x.fill_(value)


@torch.no_grad()
def normal_(x: torch.Tensor, mean: float = 0.0, std: float = 1.0) -> torch.Tensor:
    return x.normal_(mean, std)


def _fill_bias(layer, bias):
    if getattr(layer, "bias", None) is not None:
        constant_(layer.bias, bias)


@torch.no_grad()
def trunc_normal_(x: torch.Tensor, std: float = 0.02) -> torch.Tensor:
    """Normal noise truncated at two standard deviations (the reference's module-level ``TruncatedNormal(std=0.02)``)."""
    return torch.nn.init.trunc_normal_(x, 0.0, std, -2 * std, 2 * std)


def zeros_(x: torch.Tensor) -> torch.Tensor:
    return constant_(x, 0.0)


def ones_(x: torch.Tensor) -> torch.Tensor:
    return constant_(x, 1.0)


def normal_init(layer, mean: float = 0.0, std: float = 1.0, bias: float = 0.0):
    if getattr(layer, "weight", None) is not None:
        normal_(layer.weight, mean, std)
    elif isinstance(layer, torch.Tensor):           # called on a bare parameter
        normal_(layer, mean, std)
    _fill_bias(layer, bias)
    return layer


def constant_init(layer, val, bias: float = 0.0):
    if getattr(layer, "weight", None) is not None:
        constant_(layer.weight, val)
    _fill_bias(layer, bias)
    return layer

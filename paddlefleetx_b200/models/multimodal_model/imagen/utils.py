"""Small helpers of the Imagen code base under the reference's module path (ppfleetx/models/multimodal_model/imagen/utils.py:26-500): option
handling, tensor reshaping by pattern, image range conversion, masked statistics and the continuous-time noise schedule.  The models in
``modeling.py`` / ``unet.py`` inline most of this; the names exist here for code written against the reference (its T5 and DeBERTa modules
import ``rearrange`` / ``exists`` / ``default`` from this module).

The reference implements ``rearrange`` / ``repeat`` / ``reduce`` as a table of hard-coded pattern strings; here they are the general einops
operations with the reference's keyword names for axis sizes, so every pattern in that table (and any other) works.
"""
from __future__ import annotations

import math
from functools import wraps
from typing import Iterable, Optional

import einops
import torch
import torch.nn as nn
import torch.nn.functional as F

from .modeling import GaussianDiffusionContinuousTimes  # noqa: F401
from .unet import Always, Parallel, prob_mask_like, resize_image_to  # noqa: F401


# ------------------------------------------------------------------ options
def exists(val) -> bool:
    return val is not None


def identity(t, *args, **kwargs):
    return t


def first(arr, d=None):
    return arr[0] if len(arr) else d


def default(val, d):
    """``val`` unless it is None; a callable default is evaluated lazily."""
    if val is not None:
        return val
    return d() if callable(d) else d


def maybe(fn):
    @wraps(fn)
    def inner(x, *a, **k):
        return x if x is None else fn(x, *a, **k)
    return inner


def once(fn):
    state = {"done": False}

    @wraps(fn)
    def inner(*a, **k):
        if state["done"]:
            return None
        state["done"] = True
        return fn(*a, **k)
    return inner


def cast_tuple(val, length: Optional[int] = None) -> tuple:
    out = tuple(val) if isinstance(val, (list, tuple)) else (val,) * (length or 1)
    assert length is None or len(out) == length, f"expected {length} values, got {len(out)}"
    return out


def pad_tuple_to_length(t: tuple, length: int, fillvalue=None) -> tuple:
    return tuple(t) + (fillvalue,) * max(length - len(t), 0)


def is_float_dtype(dtype) -> bool:
    return dtype in (torch.float64, torch.float32, torch.float16, torch.bfloat16)


def eval_decorator(fn):
    """Run a method with the module in eval mode, then restore the mode it was in."""
    @wraps(fn)
    def inner(model, *a, **k):
        was = model.training
        model.eval()
        try:
            return fn(model, *a, **k)
        finally:
            model.train(was)
    return inner


def zero_init_(m: nn.Module) -> None:
    with torch.no_grad():
        m.weight.zero_()
        if getattr(m, "bias", None) is not None:
            m.bias.zero_()


class Identity(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, x, *a, **k):
        return x


# ------------------------------------------------------------------ tensors
def log(t: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    return torch.log(t.clamp(min=eps))


def l2norm(t: torch.Tensor) -> torch.Tensor:
    return F.normalize(t, dim=-1)


def right_pad_dims_to(x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """Append singleton dims to ``t`` until it has as many dims as ``x`` (per-sample scalars broadcast over images)."""
    return t.reshape(*t.shape, *((1,) * max(x.dim() - t.dim(), 0)))


def masked_mean(t: torch.Tensor, *, axis: int, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    if mask is None:
        return t.mean(dim=axis)
    m = mask.unsqueeze(-1).to(t.dtype)
    return (t * m).sum(dim=axis) / m.sum(dim=axis).clamp(min=1e-5)


def cast_uint8_images_to_float(images: torch.Tensor) -> torch.Tensor:
    return images.float() / 255.0 if images.dtype == torch.uint8 else images


def normalize_neg_one_to_one(img: torch.Tensor) -> torch.Tensor:
    return img * 2 - 1


def unnormalize_zero_to_one(img: torch.Tensor) -> torch.Tensor:
    return (img + 1) * 0.5


def _axes(**sizes):
    return {k: int(v) for k, v in sizes.items() if v is not None and int(v) > 0}


def rearrange(tensor, pattern: str, b: int = -1, h: int = -1, w: int = -1, c: int = -1, x: int = -1, y: int = -1, n: int = -1, s1: int = -1, s2: int = -1):
    named = _axes(b=b, h=h, w=w, c=c, x=x, y=y, n=n, s1=s1, s2=s2)
    used = {k: v for k, v in named.items() if k in pattern.replace("(", " ").replace(")", " ").split()}
    return einops.rearrange(tensor, pattern, **used)


def rearrange_many(tensors: Iterable, pattern: str, h: int = -1, x: int = -1, y: int = -1):
    return [rearrange(t, pattern, h=h, x=x, y=y) for t in tensors]


def repeat(tensor, pattern: str, h: int = -1, b: int = -1):
    named = _axes(h=h, b=b)
    used = {k: v for k, v in named.items() if k in pattern.replace("(", " ").replace(")", " ").split()}
    return einops.repeat(tensor, pattern, **used)


def repeat_many(tensors: Iterable, pattern: str, h: int = -1, b: int = -1):
    return [repeat(t, pattern, h=h, b=b) for t in tensors]


def reduce(losses, pattern: str, reduction: str = "mean"):
    return einops.reduce(losses, pattern, reduction)


class Rearrange(nn.Module):
    def __init__(self, pattern: str, **sizes):
        super().__init__()
        self.pattern, self.sizes = pattern, sizes

    def forward(self, x):
        return einops.rearrange(x, self.pattern, **self.sizes)


class EinopsToAndFrom(nn.Module):
    """Apply ``fn`` in the layout ``to_pattern`` and come back (sizes of the axes the way back needs are read from the input)."""

    def __init__(self, from_pattern: str, to_pattern: str, fn: nn.Module):
        super().__init__()
        self.from_pattern, self.to_pattern, self.fn = from_pattern, to_pattern, fn

    def forward(self, x, **kw):
        names = self.from_pattern.split()
        sizes = dict(zip(names, x.shape))
        y = einops.rearrange(x, f"{self.from_pattern} -> {self.to_pattern}")
        y = self.fn(y, **kw)
        need = {k: v for k, v in sizes.items() if k not in self.to_pattern.replace("(", " ").replace(")", " ").split() or "(" in self.to_pattern}
        return einops.rearrange(y, f"{self.to_pattern} -> {self.from_pattern}", **{k: v for k, v in need.items() if k in self.to_pattern})


# ------------------------------------------------------------------ noise schedules (continuous time, t in [0, 1])
def beta_linear_log_snr(t: torch.Tensor) -> torch.Tensor:
    return -torch.log(torch.expm1(1e-4 + 10 * t * t))


def alpha_cosine_log_snr(t: torch.Tensor, s: float = 0.008) -> torch.Tensor:
    return -log(torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** -2 - 1, eps=1e-5)


def log_snr_to_alpha_sigma(log_snr: torch.Tensor):
    return torch.sqrt(torch.sigmoid(log_snr)), torch.sqrt(torch.sigmoid(-log_snr))

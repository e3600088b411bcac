"""Small shared pieces of the folding trunk (reference ppfleetx/models/protein_folding/common.py:29-329): initialisers, chunked
evaluation for long sequences, masked means, batched gathers, the distogram featuriser and axis-shared dropout."""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
import torch.nn as nn


def no_autocast(fn: Callable) -> Callable:
    """Run ``fn`` with autocast off (cuda and cpu): frames, distances and angles are computed in the precision of their inputs (fp32), not in
    the trunk's bf16 — a batched ``rot @ vec`` would otherwise be down-cast like any other mat-mul."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        with torch.autocast("cuda", enabled=False), torch.autocast("cpu", enabled=False):
            return fn(*args, **kwargs)

    return wrapped


def set_tensor_constant(tensor: torch.Tensor, constant) -> None:
    """Fill a parameter in place (reference common.py:29-30)."""
    with torch.no_grad():
        tensor.fill_(constant)


def init_gate_linear(linear: nn.Linear) -> None:
    """Gates start open: weight 0, bias 1 (sigmoid(1) ~ 0.73)."""
    nn.init.zeros_(linear.weight)
    if linear.bias is not None:
        nn.init.ones_(linear.bias)


def init_final_linear(linear: nn.Linear) -> None:
    """Residual-branch output projections start at zero so every block is the identity at initialisation."""
    nn.init.zeros_(linear.weight)
    if linear.bias is not None:
        nn.init.zeros_(linear.bias)


def recompute_wrapper(func: Callable, *args, is_recompute: bool = True):
    if is_recompute:
        from ...parallel.recompute import recompute

        return recompute(func, *args)
    return func(*args)


def subbatch(f: Callable, arg_idx: Sequence[int], dim: Sequence[int], bs: int, out_idx: int, same_arg_idx: Optional[dict] = None) -> Callable:
    """Evaluate ``f`` in slices of size ``bs``: argument ``arg_idx[k]`` is sliced along ``dim[k]``, results are concatenated along
    ``out_idx``.  ``same_arg_idx = {a: b}`` declares that argument ``a`` is the same tensor as argument ``b`` (sliced once, passed twice).
    Bounds the attention working set for long sequences at inference time."""
    same_arg_idx = same_arg_idx or {}
    assert len(arg_idx) == len(dim)

    def wrapped(*args):
        args = list(args)
        n = args[arg_idx[0]].shape[dim[0]]
        outs = []
        for start in range(0, n, bs):
            sliced = list(args)
            for a, d in zip(arg_idx, dim):
                sliced[a] = sliced[same_arg_idx[a]] if a in same_arg_idx else args[a].narrow(d, start, min(bs, n - start))
            outs.append(f(*sliced))
        return torch.cat(outs, dim=out_idx)

    return wrapped


def batched_gather(params: torch.Tensor, indices: torch.Tensor, axis: int = 0, batch_dims: int = 0) -> torch.Tensor:
    """``params[b..., indices[b...], ...]`` along ``axis`` with ``batch_dims`` shared leading axes (tf.gather semantics)."""
    axis = axis % params.dim()
    assert axis >= batch_dims and params.shape[:batch_dims] == indices.shape[:batch_dims]
    if batch_dims == 0:
        return torch.index_select(params, axis, indices.reshape(-1)).reshape(params.shape[:axis] + indices.shape + params.shape[axis + 1:])
    idx_tail = indices.shape[batch_dims:]
    idx = indices.reshape(indices.shape[:batch_dims] + (1,) * (axis - batch_dims) + (-1,) + (1,) * (params.dim() - axis - 1))
    idx = idx.expand(params.shape[:axis] + (idx.shape[axis],) + params.shape[axis + 1:])
    out = torch.gather(params, axis, idx)
    return out.reshape(params.shape[:axis] + idx_tail + params.shape[axis + 1:])


def mask_mean(mask: torch.Tensor, value: torch.Tensor, axis=None, drop_mask_channel: bool = False, eps: float = 1e-10) -> torch.Tensor:
    """Mean of ``value`` over ``axis`` weighted by a broadcastable ``mask``."""
    if drop_mask_channel:
        mask = mask[..., 0]
    assert mask.dim() == value.dim(), "mask and value need the same rank (size-1 axes broadcast)"
    axes = list(range(value.dim())) if axis is None else ([axis] if isinstance(axis, int) else list(axis))
    mask = mask.to(value.dtype)
    bcast = 1.0
    for a in axes:
        if mask.shape[a] == 1:
            bcast *= value.shape[a]
        else:
            assert mask.shape[a] == value.shape[a]
    return (mask * value).sum(dim=axes) / (mask.sum(dim=axes) * bcast + eps)


def dgram_from_positions(positions: torch.Tensor, num_bins: int, min_bin: float, max_bin: float) -> torch.Tensor:
    """Pairwise-distance histogram features ``[..., N, N, num_bins]`` (one-hot of the squared-distance bin, last bin open-ended)."""
    lower = torch.linspace(min_bin, max_bin, num_bins, device=positions.device, dtype=positions.dtype) ** 2
    upper = torch.cat([lower[1:], lower.new_tensor([1e8])])
    d2 = ((positions.unsqueeze(-2) - positions.unsqueeze(-3)) ** 2).sum(-1, keepdim=True)
    return ((d2 > lower) & (d2 < upper)).to(positions.dtype)


class Dropout(nn.Module):
    """Dropout whose mask is shared along ``axis`` (row-wise / column-wise dropout of the Evoformer, Suppl. 1.11.6)."""

    def __init__(self, rate: float, axis=None):
        super().__init__()
        self.rate, self.axis = rate, ([axis] if isinstance(axis, int) else axis)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.training or self.rate == 0.0:
            return x
        shape = list(x.shape)
        for a in self.axis or ():
            shape[a] = 1
        keep = torch.bernoulli(torch.full(shape, 1.0 - self.rate, device=x.device, dtype=x.dtype))
        return x * keep / (1.0 - self.rate)


def __getattr__(name: str):
    # ``Transition`` lives with the Evoformer blocks here (evoformer.py); the reference keeps it in this module (common.py:189-249)
    if name == "Transition":
        from .evoformer import Transition

        return Transition
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")

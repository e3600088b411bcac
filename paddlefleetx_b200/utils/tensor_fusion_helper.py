"""Tensor-fusion façade (reference ppfleetx/utils/tensor_fusion_helper.py:23-117: ``fused_parameters`` groups params into
≤256 MB contiguous storages and ``all_reduce_parameters`` reduces the fused grads).

In this framework fusion is not an optional pass: ``parallel/flat_buffer.py`` lays every parameter and gradient out in flat
groups from optimizer construction, and the fused AdamW / ZeRO collectives / peer-memory kernels operate on those.  The
functions here expose the reference's helper API on top of that machinery for user code that called it directly.
"""
from __future__ import annotations

import torch.distributed as dist

from ..parallel.flat_buffer import FlatGroup, build_flat_groups


def fused_parameters(parameters, use_sharding=False, group_size_mb=None, pad_multiple=1):
    """Return ``(decay_fused, all_fused)`` — lists of :class:`FlatGroup`; decay membership follows the usual rule
    (no decay for 1-D tensors / ``no_weight_decay``-tagged params).  ``group_size_mb`` is accepted for API parity; a B200 has
    no reason to cap the storage so each class becomes one buffer."""
    def key(p):
        return (not getattr(p, "no_weight_decay", p.dim() <= 1),)

    groups = build_flat_groups(list(parameters), key, pad_multiple=pad_multiple)
    decay = [g for g in groups if g.key[0]]
    return decay, groups


def all_reduce_parameters(fused_groups, group=None):
    """Average the flat grad buffer of every group over ``group`` (async launches, one wait)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    works = []
    for g in fused_groups:
        buf = g.grad_buf if isinstance(g, FlatGroup) else g
        if buf is None:
            continue
        buf.div_(world)
        works.append(dist.all_reduce(buf, group=group, async_op=True))
    for w in works:
        w.wait()

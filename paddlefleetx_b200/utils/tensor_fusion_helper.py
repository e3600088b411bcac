"""Tensor-fusion façade (reference ppfleetx/utils/tensor_fusion_helper.py:23-117: ``fused_parameters`` groups params into
≤256 MB contiguous storages and ``all_reduce_parameters`` reduces the fused grads).

In this framework fusion is not an optional pass: ``parallel/flat_buffer.py`` lays every parameter and gradient out in flat
groups from optimizer construction, and the fused AdamW / ZeRO collectives / peer-memory kernels operate on those.  The
functions here expose the reference's helper API on top of that machinery for user code that called it directly.
"""
from __future__ import annotations

from collections import OrderedDict

import torch.distributed as dist

from ..parallel.flat_buffer import FlatGroup, build_flat_groups


def assign_group_by_size(parameters, group_size=256 * 1024 * 1024):
    """Greedy, order-preserving split of ``parameters`` into groups of at most ``group_size`` BYTES per dtype (a tensor larger than the cap
    gets a group of its own) — ``{group_index: [params]}`` like the reference (tensor_fusion_helper.py:30-41, Paddle's
    ``eager_assign_group_by_size``).  Pass ``group_size=None`` for one group per dtype, which is what a 180 GB part wants."""
    groups, open_group = OrderedDict(), {}            # dtype -> (index, bytes so far)
    for prm in parameters:
        size = prm.numel() * prm.element_size()
        idx, used = open_group.get(prm.dtype, (None, 0))
        if idx is None or (group_size is not None and used + size > group_size and used > 0):
            idx, used = len(groups), 0
            groups[idx] = []
        groups[idx].append(prm)
        open_group[prm.dtype] = (idx, used + size)
    return groups


def flatten_dense_tensors(parameters):
    """Move ``parameters`` (one dtype) into ONE contiguous, 256-byte-aligned storage and give them a matching flat gradient storage:
    returns the :class:`FlatGroup` twice over as ``(param_storage, grad_storage)`` — ``.param_buf`` / ``.grad_buf`` are the two flat tensors,
    every ``p.data`` and ``p.grad`` is a view into them (reference tensor_fusion_helper.py:44-81: ``ParamStorage`` / ``GradStorage``)."""
    parameters = list(parameters)
    assert parameters and all(p.requires_grad for p in parameters), "param must be trainable..."
    assert len({p.dtype for p in parameters}) == 1, "flatten_dense_tensors takes parameters of one dtype"
    (group,) = build_flat_groups(parameters, lambda p: ())
    return group, group


def obtain_storage(parameters, group_size=256 * 1024 * 1024):
    """The flat parameter buffers of ``parameters`` after grouping by size (reference tensor_fusion_helper.py:84-93)."""
    parameters = list(parameters)
    if not parameters:
        return []
    return [flatten_dense_tensors(plist)[0].param_buf for plist in assign_group_by_size(parameters, group_size).values()]


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

"""Flat parameter / gradient storage.

The reference's ``tensor_fusion`` (ppfleetx/utils/tensor_fusion_helper.py:23-117) packs parameters into
<=256 MiB, 256-byte-aligned contiguous storages so that the optimizer and the DP all-reduce touch a few big
buffers.  On a 180 GB B200 there is no reason to cap the storage: this module makes *one* contiguous buffer
per parameter class and that layout is the backbone of everything downstream —

  * the fused AdamW kernel updates a whole class in one launch,
  * ZeRO shards are plain ``[rank * S, (rank + 1) * S)`` slices of the buffer (reduce-scatter / all-gather /
    the peer-memory kernels operate on it directly, no per-tensor bookkeeping),
  * gradients accumulate in place into views of the flat grad buffer (bf16 or fp32 "main grad").

``p.data`` is re-pointed at a view of the flat buffer; ``p.grad`` (or ``p.main_grad``) is a persistent view
of the flat grad buffer.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch

ALIGN_BYTES = 256


@dataclass
class FlatGroup:
    key: Tuple
    params: List[torch.nn.Parameter]
    offsets: List[int]
    numel: int                      # padded total
    param_buf: torch.Tensor
    grad_buf: Optional[torch.Tensor] = None
    meta: Dict = field(default_factory=dict)

    def shard_range(self, rank: int, world: int) -> Tuple[int, int]:
        s = self.numel // world
        return rank * s, (rank + 1) * s

    def views(self, buf: torch.Tensor) -> List[torch.Tensor]:
        return [buf[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]


def _align(n: int, elem_bytes: int) -> int:
    a = max(ALIGN_BYTES // elem_bytes, 1)
    return ((n + a - 1) // a) * a


def build_flat_groups(params: Iterable[torch.nn.Parameter], key_fn: Callable[[torch.nn.Parameter], Tuple], pad_multiple: int = 1,
                      grad_dtype: Optional[torch.dtype] = None, allocate_grads: bool = True,
                      alloc_fn: Optional[Callable[[int, torch.dtype, torch.device], torch.Tensor]] = None) -> List[FlatGroup]:
    """Group ``params`` by ``(key_fn(p), dtype, device)`` and move each group into one contiguous buffer.

    ``pad_multiple``: the padded size is a multiple of ``pad_multiple * alignment`` so the buffer splits evenly
    into that many aligned shards.  ``alloc_fn`` lets the caller place buffers in symmetric (IPC) memory.
    """
    buckets: Dict[Tuple, List[torch.nn.Parameter]] = {}
    for p in params:
        if not p.requires_grad:
            continue
        buckets.setdefault((key_fn(p), p.dtype, p.device), []).append(p)
    groups = []
    for (key, dtype, device), plist in buckets.items():
        esize = torch.empty(0, dtype=dtype).element_size()
        offsets, cur = [], 0
        for p in plist:
            offsets.append(cur)
            cur += _align(p.numel(), esize)
        unit = max(ALIGN_BYTES // esize, 1) * max(pad_multiple, 1)
        total = ((cur + unit - 1) // unit) * unit
        alloc = alloc_fn or (lambda n, dt, dev: torch.zeros(n, dtype=dt, device=dev))
        pbuf = alloc(total, dtype, device)
        if alloc_fn is not None:
            pbuf.zero_()                 # symmetric memory is not guaranteed to arrive zeroed; the alignment padding is read by the norm / update kernels
        for p, o in zip(plist, offsets):
            view = pbuf[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        g = FlatGroup(key=key, params=plist, offsets=offsets, numel=total, param_buf=pbuf)
        if allocate_grads:
            gd = grad_dtype or dtype
            g.grad_buf = alloc(total, gd, device)
            if alloc_fn is not None:
                g.grad_buf.zero_()
            attach_grad_views(g, main_grad=(gd != dtype))
        groups.append(g)
    return groups


def attach_grad_views(g: FlatGroup, main_grad: bool) -> None:
    for p, view in zip(g.params, g.views(g.grad_buf)):
        if main_grad:
            p.main_grad = view
            p.grad = None
        else:
            p.grad = view


def zero_grads(groups: Sequence[FlatGroup]) -> None:
    for g in groups:
        if g.grad_buf is not None:
            g.grad_buf.zero_()
        for p in g.params:
            if getattr(p, "main_grad", None) is not None:
                p.grad = None

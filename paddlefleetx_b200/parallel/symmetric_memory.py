"""Symmetric (peer-mapped) device memory for the in-kernel NVLink collectives.

Every rank of a group allocates the same buffer with ``cudaMalloc`` (not the caching allocator: its blocks cannot
be exported piecewise), exports a CUDA-IPC handle, and maps all peers' handles into its own address space.  A
kernel can then load/store a peer's buffer directly — over NVLink 5 through the NVSwitch — using the pointer table
returned by ``peer_ptrs``.  A small per-rank signal pad backs ``barrier()`` (release/acquire CAS flags, see
``csrc/comm_p2p.cu``).

All methods that allocate are collective over the group (same order, same sizes on every rank).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from ..ops import _native

_SIGNAL_PAD_WORDS = 16 * 64


class SymmetricAllocator:
    def __init__(self, group):
        """``group`` is a topology ``_Group`` (ranks, process_group, rank)."""
        self.group = group
        self.world = group.nranks
        self.rank = group.rank
        self.lib = _native.require()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._allocs: List[Tuple[int, int, List[int]]] = []      # (local base, nbytes, peer bases)
        self._pad_ptrs = self._alloc_raw(_SIGNAL_PAD_WORDS * 4)[1]
        self._slot = 0

    # ---------------------------------------------------------------- allocation
    def _alloc_raw(self, nbytes: int):
        nbytes = (nbytes + 511) // 512 * 512
        ptr, handle = self.lib.ipc_alloc(nbytes)
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group.process_group)
        peers = []
        for r, h in enumerate(handles):
            peers.append(ptr if r == self.rank else self.lib.ipc_open(h))
        self._allocs.append((ptr, nbytes, peers))
        return ptr, peers

    def alloc_tensor(self, numel: int, dtype: torch.dtype, device=None) -> torch.Tensor:
        esize = torch.empty(0, dtype=dtype).element_size()
        ptr, _ = self._alloc_raw(numel * esize)
        return self.lib.tensor_from_ptr(ptr, [numel], dtype, self.device.index or 0)

    def empty(self, shape, dtype: torch.dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        return self.alloc_tensor(n, dtype).view(*shape)

    # ---------------------------------------------------------------- pointer tables
    def peer_ptrs(self, t: torch.Tensor) -> List[int]:
        p = t.data_ptr()
        for base, nbytes, peers in self._allocs:
            if base <= p < base + nbytes:
                return [pb + (p - base) for pb in peers]
        raise ValueError("tensor does not live in symmetric memory")

    def peer_tensor(self, t: torch.Tensor, peer: int) -> torch.Tensor:
        """A view of ``peer``'s copy of ``t`` (debug / tests)."""
        return self.lib.tensor_from_ptr(self.peer_ptrs(t)[peer], list(t.shape), t.dtype, self.device.index or 0)

    # ---------------------------------------------------------------- synchronisation
    def barrier(self) -> None:
        """Device-side barrier on the current stream: all ranks' prior work on their streams is visible after it."""
        self.lib.p2p_barrier(self._pad_ptrs, self.rank, self._slot)
        self._slot = (self._slot + 1) % 8


_ALLOCATORS: Dict[int, SymmetricAllocator] = {}


def get_allocator(group) -> SymmetricAllocator:
    key = id(group)
    if key not in _ALLOCATORS:
        _ALLOCATORS[key] = SymmetricAllocator(group)
    return _ALLOCATORS[key]

"""Symmetric (peer-mapped) device memory for the in-kernel NVLink collectives.

Two back-ends with one interface (``alloc_tensor`` / ``peer_ptrs`` / ``mc_ptr`` / ``barrier``):

* ``VmmSymmetricAllocator`` — the B200 path.  Arenas are physical allocations of the CUDA virtual-memory-management API
  (``cuMemCreate``, see ``csrc/symm_vmm.cpp``) exported as POSIX file descriptors, passed to the peers over unix sockets
  (``SCM_RIGHTS``) and mapped by every rank; when the fabric supports NVLink-SHARP each arena is additionally bound to a
  *multicast object*, so a kernel can reduce in the switch (``multimem.ld_reduce``) or store once to all ranks
  (``multimem.st``) through ``mc_ptr``.  Tensors are sub-allocated from 1 GiB arenas.
* ``SymmetricAllocator`` — ``cudaMalloc`` + CUDA-IPC handles, unicast only (kept for drivers / containers where VMM
  export is not available).

All methods that allocate are collective over the group (same order, same sizes on every rank).
"""
from __future__ import annotations

import os
import socket
import uuid
from typing import Dict, List, Optional, Tuple

import torch

from . import debug_poison as _debug
import torch.distributed as dist

from ..ops import _native
from ..utils.log import logger

_SIGNAL_PAD_WORDS = 16 * 64


class SymmetricAllocator:
    """cudaMalloc + CUDA-IPC back-end (unicast peer pointers only)."""

    multicast = False

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
        t = self.lib.tensor_from_ptr(ptr, [numel], dtype, self.device.index or 0)
        _debug.poison(t)                 # PFX_DEBUG_POISON=1: whatever is read before a producer wrote it is NaN, not stale data
        return t

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

    def mc_ptr(self, t: torch.Tensor) -> int:
        return 0

    def peer_tensor(self, t: torch.Tensor, peer: int) -> torch.Tensor:
        """A view of ``peer``'s copy of ``t`` (debug / tests)."""
        return self.lib.tensor_from_ptr(self.peer_ptrs(t)[peer], list(t.shape), t.dtype, self.device.index or 0)

    # ---------------------------------------------------------------- synchronisation
    def barrier(self, channel: int = 0) -> None:
        """Device-side barrier on the current stream: all ranks' prior work on their streams is visible after it.  (One channel only: use
        the VMM back-end when barriers are issued from several streams.)"""
        self.lib.p2p_barrier(self._pad_ptrs, self.rank, self._slot)
        self._slot = (self._slot + 1) % 8


class _FdExchange:
    """Moves file descriptors between the processes of a group: every rank listens on an abstract unix socket whose name
    travels through the process group; descriptors ride on ``SCM_RIGHTS`` messages."""

    def __init__(self, group):
        self.group, self.world, self.rank = group, group.nranks, group.rank
        self.name = f"\0pfx-symm-{os.getpid()}-{uuid.uuid4().hex[:12]}"
        self.listener = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self.listener.bind(self.name)
        self.listener.listen(self.world + 4)
        self.listener.settimeout(120)
        names: List[Optional[str]] = [None] * self.world
        dist.all_gather_object(names, self.name, group=group.process_group)
        self.names = names

    def _sync(self) -> None:
        tmp: List[Optional[int]] = [None] * self.world
        dist.all_gather_object(tmp, 0, group=self.group.process_group)

    def all_gather(self, fd: int) -> List[int]:
        """Every rank contributes one descriptor; returns the world's descriptors (own one at ``rank``)."""
        conns = []
        for r in range(self.world):
            if r == self.rank:
                continue
            s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            s.settimeout(120)
            s.connect(self.names[r])
            socket.send_fds(s, [self.rank.to_bytes(4, "little")], [fd])
            conns.append(s)
        out: List[int] = [-1] * self.world
        out[self.rank] = fd
        for _ in range(self.world - 1):
            c, _addr = self.listener.accept()
            c.settimeout(120)
            msg, fds, _flags, _a = socket.recv_fds(c, 4, 1)
            out[int.from_bytes(msg, "little")] = fds[0]
            c.close()
        self._sync()               # nobody closes a sending socket before its message was received
        for s in conns:
            s.close()
        return out

    def broadcast(self, fd: Optional[int], src: int) -> int:
        """``src`` hands one descriptor to every other rank."""
        conns = []
        got = fd
        if self.rank == src:
            for r in range(self.world):
                if r == src:
                    continue
                s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                s.settimeout(120)
                s.connect(self.names[r])
                socket.send_fds(s, [b"mcst"], [fd])
                conns.append(s)
        else:
            c, _addr = self.listener.accept()
            c.settimeout(120)
            _msg, fds, _flags, _a = socket.recv_fds(c, 4, 1)
            got = fds[0]
            c.close()
        self._sync()
        for s in conns:
            s.close()
        return got

    def close(self) -> None:
        self.listener.close()


class VmmSymmetricAllocator:
    """cuMem-VMM arenas mapped on every rank (+ one multicast object per arena where NVLS is available)."""

    ARENA_BYTES = 1 << 30
    ALIGN = 1024

    def __init__(self, group, want_multicast: bool = True):
        self.group = group
        self.world = group.nranks
        self.rank = group.rank
        self.lib = _native.require()
        self.device = torch.device("cuda", torch.cuda.current_device())
        caps = self.lib.vmm_caps()
        ok = bool(caps["vmm"] and caps["fd_export"])
        self._agree(ok, "CUDA VMM with POSIX-fd export is not available on every rank")
        self.multicast = bool(want_multicast and self._all(bool(caps["multicast"])) and os.environ.get("PFX_NVLS", "1") != "0")
        self._gran = int(caps["granularity"]) or (2 << 20)
        if self.multicast:
            g = int(self.lib.vmm_mc_granularity(self.world, self._gran, False))
            self._gran = max(self._gran, g or self._gran)
        self._fx = _FdExchange(group)
        self._arenas: List[dict] = []
        self._cur: Optional[dict] = None
        # flag words: [0] = multicast barrier counter, [64 + r] = unicast barrier slots
        flags = self._new_arena(self._gran, small=True)
        self._flags = flags
        self._bar_counts: Dict[int, int] = {}
        logger.info(f"symmetric memory: cuMem VMM arenas, granularity {self._gran >> 20} MiB, multicast (NVLS) {'on' if self.multicast else 'off'}")

    # ---------------------------------------------------------------- collective helpers
    def _all(self, flag: bool) -> bool:
        got: List[Optional[bool]] = [None] * self.world
        dist.all_gather_object(got, bool(flag), group=self.group.process_group)
        return all(got)

    def _agree(self, ok: bool, what: str) -> None:
        if not self._all(ok):
            raise RuntimeError(f"symmetric memory: {what}")

    # ---------------------------------------------------------------- arenas
    def _round(self, nbytes: int, small: bool) -> int:
        g = self._gran
        if self.multicast and not small:
            rec = int(self.lib.vmm_mc_granularity(self.world, max(nbytes, g), True)) or g
            if rec <= (64 << 20):          # never let a huge recommended granularity (physical memory!) blow a request up
                g = max(g, rec)
        return (nbytes + g - 1) // g * g

    def _new_arena(self, nbytes: int, small: bool = False) -> dict:
        size = self._round(nbytes, small)
        err = None
        ptr, fd = 0, -1
        try:
            ptr, fd = self.lib.vmm_arena_alloc(size)
        except RuntimeError as e:              # out of memory / driver refusal: fail on every rank, not only here
            err = e
        self._agree(err is None, f"arena allocation of {size >> 20} MiB failed on some rank ({err})")
        fds = self._fx.all_gather(int(fd))
        peers = []
        for r, f in enumerate(fds):
            peers.append(ptr if r == self.rank else int(self.lib.vmm_arena_import(int(f), size)))
        for r, f in enumerate(fds):
            os.close(f)
        arena = dict(base=int(ptr), size=size, peers=peers, mc=0, used=0)
        if self.multicast:
            arena["mc"] = self._bind_multicast(arena)
        self._arenas.append(arena)
        return arena

    def _bind_multicast(self, arena: dict) -> int:
        """create (rank 0) -> import -> add_device (all) -> barrier -> bind + map (all).  Any failure switches the whole allocator to
        unicast (collectively), it never leaves ranks disagreeing."""
        lib, size = self.lib, arena["size"]
        mc_id, fd, ok = 0, -1, True
        if self.rank == 0:
            try:
                mc_id, fd = lib.vmm_mc_create(size, self.world)
            except RuntimeError as e:
                logger.warning(f"cuMulticastCreate failed ({e}); NVLS disabled")
                ok = False
        if not self._all(ok):
            self.multicast = False
            return 0
        fd = self._fx.broadcast(int(fd) if self.rank == 0 else None, 0)
        try:
            if self.rank != 0:
                mc_id = lib.vmm_mc_import(int(fd))
            lib.vmm_mc_add_device(int(mc_id))
        except RuntimeError as e:
            logger.warning(f"multicast import / add_device failed ({e}); NVLS disabled")
            ok = False
        os.close(int(fd))
        if not self._all(ok):              # also the barrier between add_device and bind
            self.multicast = False
            return 0
        mc_ptr = 0
        try:
            mc_ptr = int(lib.vmm_mc_bind_and_map(int(mc_id), arena["base"], size))
        except RuntimeError as e:
            logger.warning(f"cuMulticastBindMem / map failed ({e}); NVLS disabled")
            ok = False
        if not self._all(ok):
            self.multicast = False
            return 0
        return mc_ptr

    # ---------------------------------------------------------------- allocation
    def alloc_tensor(self, numel: int, dtype: torch.dtype, device=None) -> torch.Tensor:
        esize = torch.empty(0, dtype=dtype).element_size()
        nbytes = (numel * esize + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        a = self._cur
        if a is None or a["used"] + nbytes > a["size"]:
            a = self._cur = self._new_arena(max(nbytes, self.ARENA_BYTES))
        ptr = a["base"] + a["used"]
        a["used"] += nbytes
        t = self.lib.tensor_from_ptr(ptr, [numel], dtype, self.device.index or 0)
        _debug.poison(t)                 # PFX_DEBUG_POISON=1: whatever is read before a producer wrote it is NaN, not stale data
        return t

    def empty(self, shape, dtype: torch.dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        return self.alloc_tensor(n, dtype).view(*shape)

    # ---------------------------------------------------------------- pointer tables
    def _find(self, p: int) -> dict:
        for a in self._arenas:
            if a["base"] <= p < a["base"] + a["size"]:
                return a
        raise ValueError("tensor does not live in symmetric memory")

    def peer_ptrs(self, t: torch.Tensor) -> List[int]:
        p = t.data_ptr()
        a = self._find(p)
        return [pb + (p - a["base"]) for pb in a["peers"]]

    def mc_ptr(self, t: torch.Tensor) -> int:
        """Multicast address of ``t`` (0 when NVLS is not available)."""
        p = t.data_ptr()
        a = self._find(p)
        return a["mc"] + (p - a["base"]) if a["mc"] else 0

    def peer_tensor(self, t: torch.Tensor, peer: int) -> torch.Tensor:
        return self.lib.tensor_from_ptr(self.peer_ptrs(t)[peer], list(t.shape), t.dtype, self.device.index or 0)

    # ---------------------------------------------------------------- synchronisation
    def barrier(self, channel: int = 0) -> None:
        """Device-side barrier on the current stream (release / acquire at system scope): everything the ranks wrote before it
        — locally, to peers or through the switch — is visible to every rank's kernels after it.  Barriers issued from DIFFERENT
        streams must use different ``channel`` s (0..7): each channel has its own flag words and arrival counter, so two streams can
        never satisfy each other's barrier, whatever order the ranks' schedulers run them in."""
        assert 0 <= channel < 8
        cnt = self._bar_counts[channel] = self._bar_counts.get(channel, 0) + 1
        f = self._flags
        off = channel * 1024                                   # bytes: [0, 256) multicast counter word, [256, 1024) unicast slots
        if self.multicast and f["mc"]:
            self.lib.nvls_barrier(f["mc"] + off, f["base"] + off, (cnt * self.world) & 0xFFFFFFFF)
        else:
            self.lib.p2p_flag_barrier([p + off + 256 for p in f["peers"]], self.rank, cnt & 0xFFFFFFFF)


_ALLOCATORS: Dict[int, object] = {}


def get_allocator(group):
    """The group's allocator: cuMem-VMM (+ NVLS) when the driver offers it, CUDA-IPC otherwise (``PFX_SYMM_BACKEND=ipc`` forces it)."""
    key = id(group)
    if key not in _ALLOCATORS:
        backend = os.environ.get("PFX_SYMM_BACKEND", "auto")
        alloc = None
        if backend != "ipc":
            caps = _native.require().vmm_caps()
            got: List[Optional[bool]] = [None] * group.nranks
            dist.all_gather_object(got, bool(caps["vmm"] and caps["fd_export"]), group=group.process_group)
            if all(got):
                alloc = VmmSymmetricAllocator(group)
            elif backend == "vmm":
                raise RuntimeError("PFX_SYMM_BACKEND=vmm but CUDA VMM / fd export is not available on every rank")
        _ALLOCATORS[key] = alloc if alloc is not None else SymmetricAllocator(group)
    return _ALLOCATORS[key]

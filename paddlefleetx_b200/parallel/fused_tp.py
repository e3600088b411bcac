"""Fused compute+collective linears for tensor/sequence parallelism (NVSwitch boxes).

    all_gather_linear(x[s/n,b,h], W)      = ONE kernel: peers push their sequence shard into this rank's gather buffer while
                                            the tcgen05 mainloop already consumes the local shard (csrc/gemm_sm100.cu, AG mode)
    linear_reduce_scatter(x[s,b,in/n], W) = GEMM whose epilogue stores every tile into the owner rank's staging slot
                                            over NVLink, followed by a local slot reduction

They replace the ``all-gather -> GEMM`` / ``GEMM -> reduce-scatter`` pairs of the reference's sequence-parallel
linears (gpt/dygraph/sequence_parallel_utils.py:215-398, NCCL call + cuBLAS call).  Backward passes use the mirrored
fused kernel for the activation gradient (AG-GEMM <-> GEMM-RS) and plain GEMMs on the already gathered operand for
the weight gradient.  Buffers live in symmetric memory and are cached per (group, shape).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from ..ops import _native
from ..ops import functional as OF
from . import comm_ops as C
from .symmetric_memory import get_allocator

_MAX_COMM_CTAS = 32
_epoch = [0]
_bufs: Dict[Tuple, dict] = {}


def _ag_buffers(group, rows: int, K: int, dtype):
    """Two alternating gather buffers (+ flag arrays) per shape: a peer may already push call k+1 while this rank still
    reads call k; it can never be two calls ahead because its kernel k needs this rank's pushes of call k to finish."""
    key = ("ag", id(group), rows, K, dtype)
    if key not in _bufs:
        sm = get_allocator(group)
        chunk = 256
        sets = []
        for _ in range(2):
            gathered = sm.empty((rows * group.nranks, K), dtype)
            flags = sm.alloc_tensor(group.nranks * max(rows // chunk, 1) + 64, torch.int32)
            flags.zero_()
            sets.append(dict(gathered=gathered, flags=flags, peer_gather=sm.peer_ptrs(gathered), peer_flags=sm.peer_ptrs(flags)))
        torch.cuda.synchronize()
        sm.barrier()
        _bufs[key] = dict(sm=sm, sets=sets, chunk=chunk, calls=0)
    return _bufs[key]


def _rs_buffers(group, rows_per_rank: int, N: int, dtype):
    """Two alternating staging buffers: the single barrier between scatter and reduce of call k also proves that every
    rank finished reducing call k-1, so slot set (k+1) % 2 is free when call k+1 starts."""
    key = ("rs", id(group), rows_per_rank, N, dtype)
    if key not in _bufs:
        sm = get_allocator(group)
        sets = []
        for _ in range(2):
            staging = sm.empty((group.nranks, rows_per_rank, N), dtype)
            sets.append(dict(staging=staging, peer_staging=sm.peer_ptrs(staging)))
        torch.cuda.synchronize()
        sm.barrier()
        _bufs[key] = dict(sm=sm, sets=sets, calls=0)
    return _bufs[key]


def fused_ok(x2: torch.Tensor, group, rows_per_rank: int) -> bool:
    return (x2.is_cuda and x2.dtype == torch.bfloat16 and group is not None and 2 <= group.nranks <= 8 and rows_per_rank % 256 == 0
            and _native.available())


def ag_gemm(x_shard: torch.Tensor, w: torch.Tensor, bias, group, b_kmajor: bool = True):
    """gathered(x_shard)[M,K] @ op(w) -> ([M,N], gathered view).  ``w`` is [N,K] (b_kmajor) or [K,N]."""
    lib = _native.require()
    rows, K = x_shard.shape
    b = _ag_buffers(group, rows, K, x_shard.dtype)
    cur = b["sets"][b["calls"] % 2]
    b["calls"] += 1
    _epoch[0] += 1
    items = max(rows // b["chunk"], 1) * (group.nranks - 1)          # (chunk, peer) push items, see ag_push_role
    comm_ctas = max(2, min(_MAX_COMM_CTAS, items) // 2 * 2)          # even: the GEMM runs CTA pairs
    y = lib.gemm_ag(x_shard.contiguous(), w, cur["gathered"], cur["peer_gather"], cur["peer_flags"], cur["flags"], bias, group.rank,
                    b["chunk"], comm_ctas, _epoch[0], b_kmajor, 0)
    OF._count()
    # the local shard is consumed straight from x_shard by the kernel; complete the gathered buffer for later (wgrad) use
    cur["gathered"][group.rank * rows:(group.rank + 1) * rows].copy_(x_shard)
    return y, cur["gathered"]


def gemm_rs(a: torch.Tensor, w: torch.Tensor, group, a_kmajor: bool = True, b_kmajor: bool = True, bias=None) -> torch.Tensor:
    """reduce_scatter_rows(a @ op(w)) -> [M/n, N]."""
    lib = _native.require()
    M = a.shape[0] if a_kmajor else a.shape[1]
    N = w.shape[0] if b_kmajor else w.shape[1]
    rpr = M // group.nranks
    b = _rs_buffers(group, rpr, N, a.dtype)
    cur = b["sets"][b["calls"] % 2]
    b["calls"] += 1
    lib.gemm_rs_scatter(a, w, cur["peer_staging"], group.rank, rpr, a_kmajor, b_kmajor, 0)
    b["sm"].barrier()                        # every rank's tiles have landed (and everyone is done with the other slot set)
    out = lib.slot_reduce(cur["staging"], bias, group.nranks)
    OF._count(3)
    return out


_side_stream = [None]


def _regather_async(x_shard: torch.Tensor, group):
    """All-gather ``x_shard`` [rows, K] over the group with OUR kernel (multimem / peer stores) on a side stream; returns the gathered
    [rows * n, K] buffer (symmetric, two alternating per shape) and the event the consumer has to wait for.  Barrier channel 1: the main
    stream's fused kernels synchronise on channel 0 at the same time."""
    lib = _native.require()
    rows, K = x_shard.shape
    key = ("regather", id(group), rows, K, x_shard.dtype)
    if key not in _bufs:
        sm = get_allocator(group)
        sets = [sm.empty((rows * group.nranks, K), x_shard.dtype) for _ in range(2)]
        torch.cuda.synchronize()
        sm.barrier()
        _bufs[key] = dict(sm=sm, sets=sets, calls=0)
    b = _bufs[key]
    buf = b["sets"][b["calls"] % 2]
    b["calls"] += 1
    sm = b["sm"]
    if _side_stream[0] is None:
        _side_stream[0] = torch.cuda.Stream()
    side, cur = _side_stream[0], torch.cuda.current_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        sm.barrier(channel=1)        # every rank is past the previous use of this buffer set (two calls ago) and x_shard is final
        lib.symm_all_gather(sm.mc_ptr(buf) if group.nranks >= 4 else 0, sm.peer_ptrs(buf), group.rank * rows * K * x_shard.element_size(), x_shard, group.rank, 64)
        sm.barrier(channel=1)        # all shards have landed
        ev = torch.cuda.Event()
        ev.record(side)
    x_shard.record_stream(side)
    OF._count(3)
    return buf, ev


class _AllGatherLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, group):
        s_local, bsz, h = x.shape
        x2 = x.reshape(s_local * bsz, h)
        y, _ = ag_gemm(x2, weight, bias, group)
        ctx.save_for_backward(x2, weight)
        ctx.group, ctx.shape, ctx.has_bias = group, (s_local, bsz, h), bias is not None
        return y.view(s_local * group.nranks, bsz, weight.shape[0])

    @staticmethod
    def backward(ctx, gy):
        lib = _native.require()
        x2, weight = ctx.saved_tensors
        group = ctx.group
        s_local, bsz, h = ctx.shape
        g2 = gy.reshape(-1, gy.shape[-1]).contiguous()
        # dW = dY^T @ X_full needs the gathered input again (the forward's gather buffer has been reused by later layers): our all-gather
        # kernel re-gathers it on the side stream, underneath the dgrad GEMM -> reduce-scatter below
        x_full, ready = _regather_async(x2, group)
        # dX shard = reduce_scatter(dY @ W): fused GEMM -> RS (B operand MN-major, no transpose copy)
        gx = gemm_rs(g2, weight, group, True, False).view(s_local, bsz, h)
        torch.cuda.current_stream().wait_event(ready)
        gw = OF._wgrad(lib, g2, x_full, weight)          # straight into the flat optimizer's main_grad when there is one
        gb = lib.colsum(g2, False) if ctx.has_bias else None
        OF._count(2)
        return gx, gw, gb, None


class _LinearReduceScatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, group):
        s, bsz, k = x.shape
        x2 = x.reshape(s * bsz, k)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        out = gemm_rs(x2, weight, group)
        ctx.save_for_backward(x2, weight)
        ctx.group, ctx.shape = group, (s, bsz, k)
        return out.view(s // group.nranks, bsz, weight.shape[0])

    @staticmethod
    def backward(ctx, gy):
        lib = _native.require()
        x2, weight = ctx.saved_tensors
        group = ctx.group
        s, bsz, k = ctx.shape
        g2 = gy.reshape(-1, gy.shape[-1]).contiguous()
        # dX = all_gather(dY) @ W: fused AG -> GEMM; the gathered dY is then reused for the weight gradient
        gx, g_full = ag_gemm(g2, weight, None, group, b_kmajor=False)
        gw = OF._wgrad(lib, g_full, x2, weight)
        return gx.view(s, bsz, k), gw, None


def all_gather_linear(x: torch.Tensor, weight: torch.Tensor, bias, group) -> torch.Tensor:
    """x: [s/n, b, h] (sequence shard) -> [s, b, out/n]."""
    rows = x.shape[0] * x.shape[1]
    if not fused_ok(x, group, rows):
        return OF.linear(C.all_gather_seq(x, group), weight, bias)
    # rows of the gathered matrix must be rank-contiguous: [s/n, b, h] flattened is (s_local, b) row-major per rank -> OK
    return _AllGatherLinear.apply(x.contiguous(), weight, bias, group)


def linear_reduce_scatter(x: torch.Tensor, weight: torch.Tensor, group) -> torch.Tensor:
    """x: [s, b, in/n] -> [s/n, b, out] (summed over the group)."""
    rows = x.shape[0] * x.shape[1]
    if not fused_ok(x, group, rows // group.nranks):
        return C.reduce_scatter_seq(OF.linear(x, weight, None), group)
    return _LinearReduceScatter.apply(x, weight, group)

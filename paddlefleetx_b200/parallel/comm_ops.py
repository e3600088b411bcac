"""Autograd-aware collectives for tensor / sequence parallelism.

These are the primitives the reference reaches through ``collective._c_identity`` / ``_c_concat`` /
``_mp_allreduce`` (gpt/dygraph/hybrid_model.py:75-84) and through its in-tree sequence-parallel
PyLayers ``ScatterOp / GatherOp / AllGatherOp / ReduceScatterOp``
(gpt/dygraph/sequence_parallel_utils.py:41-137).  Every op is a fwd/bwd conjugate pair:

    identity      <-> all-reduce          (copy_to_group)
    all-reduce    <-> identity            (reduce_from_group)
    all-gather    <-> reduce-scatter      (all_gather_seq / along last dim: gather <-> split)
    reduce-scatter<-> all-gather          (reduce_scatter_seq)
    split         <-> all-gather          (scatter_seq)

A ``group`` is the ``_Group`` wrapper from ``topology.py`` (or None / size 1 = no-op).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def group_size(group) -> int:
    return 1 if group is None else group.nranks


def group_rank(group) -> int:
    return 0 if group is None else max(group.rank, 0)


def _pg(group):
    return None if group is None else group.process_group


def _active(group) -> bool:
    return group is not None and group.nranks > 1 and group.process_group is not None


# ------------------------------------------------------------------ raw (non-autograd) helpers
def all_reduce_(t: torch.Tensor, group, op=dist.ReduceOp.SUM) -> torch.Tensor:
    if _active(group):
        dist.all_reduce(t, op=op, group=_pg(group))
    return t


def all_gather_dim0(t: torch.Tensor, group) -> torch.Tensor:
    if not _active(group):
        return t
    t = t.contiguous()
    out = torch.empty((group.nranks * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t, group=_pg(group))
    return out


def reduce_scatter_dim0(t: torch.Tensor, group) -> torch.Tensor:
    if not _active(group):
        return t
    t = t.contiguous()
    assert t.shape[0] % group.nranks == 0, f"dim0 {t.shape[0]} not divisible by {group.nranks}"
    out = torch.empty((t.shape[0] // group.nranks,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    if t.device.type == "cpu":
        # gloo has no reduce_scatter: all-reduce then slice (CPU correctness path only)
        full = t.clone()
        dist.all_reduce(full, group=_pg(group))
        out.copy_(full.chunk(group.nranks, dim=0)[group.rank])
    else:
        dist.reduce_scatter_tensor(out, t, group=_pg(group))
    return out


def split_dim(t: torch.Tensor, group, dim: int) -> torch.Tensor:
    if group_size(group) == 1:
        return t
    return t.chunk(group.nranks, dim=dim)[group.rank].contiguous()


def all_gather_dim(t: torch.Tensor, group, dim: int) -> torch.Tensor:
    if not _active(group):
        return t
    if dim == 0:
        return all_gather_dim0(t, group)
    moved = t.movedim(dim, 0)
    return all_gather_dim0(moved, group).movedim(0, dim).contiguous()


# ------------------------------------------------------------------------- autograd pairs
class _CopyToGroup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x

    @staticmethod
    def backward(ctx, g):
        return all_reduce_(g.contiguous().clone() if not g.is_contiguous() else g.clone(), ctx.group), None


class _ReduceFromGroup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        return all_reduce_(x.contiguous().clone(), group)

    @staticmethod
    def backward(ctx, g):
        return g, None


class _GatherLastDim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return all_gather_dim(x, group, x.dim() - 1)

    @staticmethod
    def backward(ctx, g):
        return split_dim(g, ctx.group, g.dim() - 1), None


class _ScatterLastDim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return split_dim(x, group, x.dim() - 1)

    @staticmethod
    def backward(ctx, g):
        return all_gather_dim(g, ctx.group, g.dim() - 1), None


class _ScatterSeq(torch.autograd.Function):
    """split along dim 0 (sequence) fwd / all-gather bwd — reference ``ScatterOp``."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return split_dim(x, group, 0)

    @staticmethod
    def backward(ctx, g):
        return all_gather_dim0(g, ctx.group), None


class _GatherSeq(torch.autograd.Function):
    """all-gather along dim 0 fwd / split bwd — reference ``GatherOp``."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return all_gather_dim0(x, group)

    @staticmethod
    def backward(ctx, g):
        return split_dim(g, ctx.group, 0), None


class _AllGatherSeq(torch.autograd.Function):
    """all-gather fwd / reduce-scatter bwd — reference ``AllGatherOp``."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return all_gather_dim0(x, group)

    @staticmethod
    def backward(ctx, g):
        return reduce_scatter_dim0(g, ctx.group), None


class _ReduceScatterSeq(torch.autograd.Function):
    """reduce-scatter fwd / all-gather bwd — reference ``ReduceScatterOp``."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return reduce_scatter_dim0(x, group)

    @staticmethod
    def backward(ctx, g):
        return all_gather_dim0(g, ctx.group), None


def copy_to_group(x, group):
    return x if not _active(group) else _CopyToGroup.apply(x, group)


def reduce_from_group(x, group):
    return x if not _active(group) else _ReduceFromGroup.apply(x, group)


def gather_last_dim(x, group):
    return x if not _active(group) else _GatherLastDim.apply(x, group)


def scatter_last_dim(x, group):
    return x if group_size(group) == 1 else _ScatterLastDim.apply(x, group)


def scatter_seq(x, group):
    return x if group_size(group) == 1 else _ScatterSeq.apply(x, group)


def gather_seq(x, group):
    return x if not _active(group) else _GatherSeq.apply(x, group)


def all_gather_seq(x, group):
    return x if not _active(group) else _AllGatherSeq.apply(x, group)


def reduce_scatter_seq(x, group):
    return x if not _active(group) else _ReduceScatterSeq.apply(x, group)


# reference-compatible names (hybrid_model.py:75-84)
_c_identity = copy_to_group
_mp_allreduce = reduce_from_group
_c_concat = gather_last_dim
_c_split = scatter_last_dim


def broadcast_params(module: torch.nn.Module, group, src_rank: int, skip_distributed: bool = False) -> None:
    """``sync_params_buffers``: broadcast params+buffers from the group's src rank."""
    if not _active(group):
        return
    for t in list(module.parameters()) + list(module.buffers()):
        if skip_distributed and getattr(t, "tp_sharded", False):
            continue
        if getattr(t, "no_sync", False):          # expert parameters differ per rank by design
            continue
        dist.broadcast(t.data, src=src_rank, group=_pg(group))


_SIDE_STREAMS: list = []      # communication streams whose in-flight gradient reductions touch the same buffers (optimizer bucket reducers)


def register_side_stream(stream) -> None:
    if stream is not None and stream not in _SIDE_STREAMS:
        _SIDE_STREAMS.append(stream)


def wait_side_streams() -> None:
    """Order the current stream after every registered communication stream.  Post-backward collectives on gradient buffers
    (tied-embedding all-reduce, sequence-parallel gradient all-reduce) call this first: with ``reduce_overlap`` the bucket
    reduce-scatter / all-reduce of the same buffer may still be running on the side stream."""
    if _SIDE_STREAMS and torch.cuda.is_available():
        cur = torch.cuda.current_stream()
        for s in _SIDE_STREAMS:
            cur.wait_stream(s)


_PRE_BACKWARD: list = []      # weak references to bound methods (an optimizer must not be kept alive by this registry)


def register_pre_backward(fn) -> None:
    """Callbacks run right before any backward pass starts writing gradients (the engine and the pipeline runtime call
    ``run_pre_backward``): the overlapped optimizer uses it to make sure no side-stream update still reads the gradient buffers."""
    import weakref

    _PRE_BACKWARD.append(weakref.WeakMethod(fn) if hasattr(fn, "__self__") else (lambda fn=fn: fn))


def run_pre_backward() -> None:
    dead = False
    for ref in _PRE_BACKWARD:
        fn = ref()
        if fn is None:
            dead = True
        else:
            fn()
    if dead:
        _PRE_BACKWARD[:] = [r for r in _PRE_BACKWARD if r() is not None]


def fused_allreduce_gradients(params, group, scale: Optional[float] = None) -> None:
    """Coalesced grad all-reduce (÷ nranks by default) — reference
    ``fused_allreduce_gradients[_with_group]`` (eager_engine.py:491-504)."""
    if not _active(group):
        return
    wait_side_streams()
    grads = [p.grad if getattr(p, "main_grad", None) is None else p.main_grad for p in params]
    grads = [g for g in grads if g is not None]
    if not grads:
        return
    scale = 1.0 / group.nranks if scale is None else scale
    by_dtype = {}
    for g in grads:
        by_dtype.setdefault(g.dtype, []).append(g)
    for dtype, gs in by_dtype.items():
        flat = torch._utils._flatten_dense_tensors(gs)
        if scale != 1.0:
            flat.mul_(scale)
        dist.all_reduce(flat, group=_pg(group))
        for g, synced in zip(gs, torch._utils._unflatten_dense_tensors(flat, gs)):
            g.copy_(synced)


# ------------------------------------------------------------------------------------------------ context parallelism (Ulysses)
def _all_to_all_dims(x: torch.Tensor, group, scatter_dim: int, gather_dim: int) -> torch.Tensor:
    """Split ``scatter_dim`` over the group, concatenate what arrives (rank order) along ``gather_dim``."""
    w = group_size(group)
    if w == 1 or group.process_group is None:
        return x
    parts = [p.contiguous() for p in x.chunk(w, dim=scatter_dim)]
    out = [torch.empty_like(parts[0]) for _ in range(w)]
    if x.is_cuda:
        dist.all_to_all(out, parts, group=group.process_group)
    else:               # gloo has no all_to_all: pairwise exchange
        reqs = [dist.isend(parts[r], group.ranks[r], group=group.process_group) for r in range(w) if r != group.rank]
        for r in range(w):
            if r == group.rank:
                out[r].copy_(parts[r])
            else:
                dist.recv(out[r], group.ranks[r], group=group.process_group)
        for q in reqs:
            q.wait()
    return torch.cat(out, dim=gather_dim)


class _SeqHeadAllToAll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group, scatter_dim, gather_dim):
        ctx.args = (group, scatter_dim, gather_dim)
        return _all_to_all_dims(x, group, scatter_dim, gather_dim)

    @staticmethod
    def backward(ctx, g):
        group, s, ga = ctx.args
        return _all_to_all_dims(g.contiguous(), group, ga, s), None, None, None


def seq_head_all_to_all(x: torch.Tensor, group, scatter_dim: int, gather_dim: int) -> torch.Tensor:
    """The Ulysses exchange around attention: ``[b, s/c, H, d] -> [b, s, H/c, d]`` (scatter heads, gather sequence) and back.  Autograd-aware:
    the backward is the inverse exchange."""
    return x if group_size(group) == 1 else _SeqHeadAllToAll.apply(x, group, scatter_dim, gather_dim)


"""Dynamic Axial Parallelism (FastFold-style) primitives — reference distributed/protein_folding/dap.py:74-426.

MSA / pair activations are sharded along one of their two "sequence" axes across the ``dap`` group; ``row_to_col`` /
``col_to_row`` transpose which axis is sharded with an **all-to-all** (the Ulysses-like operation of this code base);
``scatter`` / ``gather`` / ``all_gather`` move between sharded and replicated layouts; gradients of replicated parameters
are all-reduced over the group.  Every op is an autograd function with the conjugate collective in backward.  The
reference declares async "duality" variants but implements them synchronously (dap.py:81-82,163-164); here the ``*_opp``
names are provided as synchronous aliases as well.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ...parallel import comm_ops as C
from .scg import scg


def _grp(group=None):
    return group if group is not None else scg.get_dap_group()


def _all_to_all(x: torch.Tensor, group, scatter_dim: int, gather_dim: int) -> torch.Tensor:
    """Split ``scatter_dim`` over the group, concatenate received pieces along ``gather_dim``."""
    w = C.group_size(group)
    if w == 1 or group.process_group is None:
        return x
    parts = [p.contiguous() for p in x.chunk(w, dim=scatter_dim)]
    out = [torch.empty_like(parts[0]) for _ in range(w)]
    if x.is_cuda:
        dist.all_to_all(out, parts, group=group.process_group)
    else:   # gloo: emulate
        reqs = [dist.isend(parts[r], group.ranks[r], group=group.process_group) for r in range(w) if r != group.rank]
        for r in range(w):
            if r == group.rank:
                out[r].copy_(parts[r])
            else:
                dist.recv(out[r], group.ranks[r], group=group.process_group)
        for q in reqs:
            q.wait()
    return torch.cat(out, dim=gather_dim)


class _Scatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dim, group):
        ctx.dim, ctx.group = dim, group
        return C.split_dim(x, group, dim)

    @staticmethod
    def backward(ctx, g):
        return C.all_gather_dim(g.contiguous(), ctx.group, ctx.dim), None, None


class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dim, group):
        ctx.dim, ctx.group = dim, group
        return C.all_gather_dim(x.contiguous(), group, dim)

    @staticmethod
    def backward(ctx, g):
        return C.split_dim(g, ctx.group, ctx.dim), None, None


class _AllGatherSum(torch.autograd.Function):
    """all-gather whose backward is reduce-scatter (used when the gathered tensor feeds a computation every rank repeats)."""

    @staticmethod
    def forward(ctx, x, dim, group):
        ctx.dim, ctx.group = dim, group
        return C.all_gather_dim(x.contiguous(), group, dim)

    @staticmethod
    def backward(ctx, g):
        grp = ctx.group
        if C.group_size(grp) == 1 or grp.process_group is None:
            return g, None, None
        g = g.contiguous().clone()
        dist.all_reduce(g, group=grp.process_group)
        return C.split_dim(g, grp, ctx.dim), None, None


class _AllToAll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scatter_dim, gather_dim, group):
        ctx.args = (scatter_dim, gather_dim, group)
        return _all_to_all(x, group, scatter_dim, gather_dim)

    @staticmethod
    def backward(ctx, g):
        s, ga, grp = ctx.args
        return _all_to_all(g.contiguous(), grp, ga, s), None, None, None


def scatter(x, axis: int = 0, group=None):
    g = _grp(group)
    return x if C.group_size(g) == 1 else _Scatter.apply(x, axis, g)


def gather(x, axis: int = 0, group=None):
    g = _grp(group)
    return x if C.group_size(g) == 1 else _Gather.apply(x, axis, g)


def all_gather(x, axis: int = 0, group=None):
    g = _grp(group)
    return x if C.group_size(g) == 1 else _AllGatherSum.apply(x, axis, g)


def all_to_all(x, in_axis: int, out_axis: int, group=None):
    g = _grp(group)
    return x if C.group_size(g) == 1 else _AllToAll.apply(x, in_axis, out_axis, g)


def row_to_col(x, group=None):
    """[.., R/n, C, ..] (rows sharded, dims 1/2) -> [.., R, C/n, ..] (columns sharded)."""
    return all_to_all(x, 2, 1, group)


def col_to_row(x, group=None):
    return all_to_all(x, 1, 2, group)


# "duality async" names of the reference (synchronous there as well)
all_gather_opp = all_gather
all_to_all_opp = all_to_all


def grad_sync(params, group=None) -> None:
    """All-reduce (mean-free SUM) the gradients of parameters replicated across the dap group (dap.py tail)."""
    C.fused_allreduce_gradients(list(params), _grp(group), scale=1.0)

"""Branch Parallelism (degree 2) — reference distributed/protein_folding/bp.py:38-152: inside one Evoformer block the MSA
branch runs on bp rank 0 and the pair branch on bp rank 1; ``broadcast`` shares each branch's output with the other rank
(backward broadcasts/reduces the gradient back to the producer), ``all_reduce`` sums replicated-parameter grads."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ...parallel import comm_ops as C
from .scg import scg


def _grp(group=None):
    return group if group is not None else scg.get_bp_group()


def get_world_size() -> int:
    return scg.get_bp_world_size()


def get_rank_in_group() -> int:
    return scg.get_bp_rank()


class _BroadcastGrad(torch.autograd.Function):
    """fwd: broadcast from ``src`` (index in group); bwd: sum grads onto the producer, zero elsewhere."""

    @staticmethod
    def forward(ctx, x, src, group):
        ctx.src, ctx.group = src, group
        out = x.contiguous().clone()
        dist.broadcast(out, src=group.ranks[src], group=group.process_group)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, group=ctx.group.process_group)
        if ctx.group.rank != ctx.src:
            g = torch.zeros_like(g)
        return g, None, None


class _GradBroadcast(torch.autograd.Function):
    """fwd: identity; bwd: every rank receives ``src``'s gradient (reference bp.py ``broadcast_grad_for_backward``)."""

    @staticmethod
    def forward(ctx, x, src, group):
        ctx.src, ctx.group = src, group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.broadcast(g, src=ctx.group.ranks[ctx.src], group=ctx.group.process_group)
        return g, None, None


class BroadcastGrad(torch.autograd.Function):
    """Reference class form (bp.py:51-62): ``BroadcastGrad.apply(input, src)`` — identity forward, the gradient of rank ``src`` of the bp
    group reaches every rank in backward."""

    @staticmethod
    def forward(ctx, input, src):
        ctx.src = src
        return input.clone()

    @staticmethod
    def backward(ctx, grad_output):
        g, grp = grad_output.contiguous().clone(), _grp()
        if C.group_size(grp) > 1 and grp.process_group is not None:
            dist.broadcast(g, src=grp.ranks[ctx.src], group=grp.process_group)
        return g, None


def broadcast_grad_for_backward(x, src: int = 0, group=None):
    g = _grp(group)
    if C.group_size(g) == 1 or g.process_group is None:
        return x
    return _GradBroadcast.apply(x, src, g)


class _ReplicatedExit(torch.autograd.Function):
    """fwd: identity; bwd: grad / group size.  Placed where the branch-parallel trunk hands its outputs to code every rank repeats (heads,
    loss): each rank then back-propagates the same gradient, and the last ``broadcast`` sums them onto the producer — once is right."""

    @staticmethod
    def forward(ctx, x, n):
        ctx.n = n
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g / ctx.n, None


def replicated_exit(x, group=None):
    g = _grp(group)
    n = C.group_size(g)
    return x if n == 1 or g.process_group is None else _ReplicatedExit.apply(x, n)


def broadcast(x, src: int = 0, group=None):
    g = _grp(group)
    if C.group_size(g) == 1 or g.process_group is None:
        return x
    return _BroadcastGrad.apply(x, src, g)


def all_reduce(x, group=None):
    return C.reduce_from_group(x, _grp(group))


class _SyncEvoformerResults(torch.autograd.Function):
    """One node for the block's three exchanges, so both ranks issue the backward collectives in the same order whatever order the autograd
    engine would pick for independent nodes (the two ranks hold different graphs).

    fwd: msa <- rank 0, pair <- rank 1, outer <- rank 0; returns (msa, pair + outer).
    bwd: the consumers of the outputs are split across the ranks, so their gradients are summed; the sum goes to each tensor's producer."""

    @staticmethod
    def forward(ctx, msa, pair, outer, group):
        ctx.group = group
        pg, ranks = group.process_group, group.ranks
        msa, pair, outer = msa.contiguous().clone(), pair.contiguous().clone(), outer.contiguous().clone()
        dist.broadcast(msa, src=ranks[0], group=pg)
        dist.broadcast(pair, src=ranks[1], group=pg)
        dist.broadcast(outer, src=ranks[0], group=pg)
        return msa, pair + outer

    @staticmethod
    def backward(ctx, g_msa, g_pair):
        pg, rank = ctx.group.process_group, ctx.group.rank
        g_msa, g_pair = g_msa.contiguous().clone(), g_pair.contiguous().clone()
        dist.all_reduce(g_msa, group=pg)
        dist.all_reduce(g_pair, group=pg)
        zero_m, zero_p = torch.zeros_like(g_msa), torch.zeros_like(g_pair)
        return (g_msa if rank == 0 else zero_m), (g_pair if rank == 1 else zero_p), (g_pair if rank == 0 else zero_p), None


class SyncEvoformerResults(torch.autograd.Function):
    """Reference class form and argument order (bp.py:90-111): ``msa, pair = SyncEvoformerResults.apply(outer, msa, pair)`` on the bp group."""

    @staticmethod
    def forward(ctx, outer, msa, pair):
        ctx.group = _grp()
        return _SyncEvoformerResults.forward(ctx, msa, pair, outer, ctx.group)

    @staticmethod
    def backward(ctx, g_msa, g_pair):
        gm, gp, go, _ = _SyncEvoformerResults.backward(ctx, g_msa, g_pair)
        return go, gm, gp


def sync_evoformer_results(msa_act, pair_act, outer=None, group=None):
    """``SyncEvoformerResults`` (reference bp.py:84-113): rank 0 owns the fresh MSA activation and the outer-product-mean update, rank 1 the
    fresh pair activation; every rank leaves with ``(msa, pair + outer)``."""
    g = _grp(group)
    if outer is None:
        outer = torch.zeros_like(pair_act)
    if C.group_size(g) == 1 or g.process_group is None:
        return msa_act, pair_act + outer
    return _SyncEvoformerResults.apply(msa_act, pair_act, outer, g)


def grad_sync(params, group=None) -> None:
    """Sum the gradients of parameters replicated across the bp group.  Accepts a parameter list or the reference's optimizer
    ``param_groups`` (only groups flagged ``bp: True`` take part; bp.py:127-152)."""
    params = list(params)
    if params and isinstance(params[0], dict):
        params = [q for grp in params if grp.get("bp", False) for q in grp["params"] if not getattr(q, "tp_sharded", False)]
    C.fused_allreduce_gradients(params, _grp(group), scale=1.0)

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


def broadcast(x, src: int = 0, group=None):
    g = _grp(group)
    if C.group_size(g) == 1 or g.process_group is None:
        return x
    return _BroadcastGrad.apply(x, src, g)


def all_reduce(x, group=None):
    return C.reduce_from_group(x, _grp(group))


def sync_evoformer_results(msa_act, pair_act, group=None):
    """``SyncEvoformerResults``: rank 0 owns the fresh MSA activation, rank 1 the fresh pair activation."""
    return broadcast(msa_act, 0, group), broadcast(pair_act, 1, group)


def grad_sync(params, group=None) -> None:
    C.fused_allreduce_gradients(list(params), _grp(group), scale=1.0)

"""Data-parallel helpers of the protein-folding stack (reference distributed/protein_folding/dp.py:28-107): broadcast
parameters from the dp source, all-reduce (+ average) every gradient after backward, all-reduce a metric tensor."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ...parallel import comm_ops as C
from .scg import scg


def get_world_size() -> int:
    return scg.get_dp_world_size()


def get_rank_in_group() -> int:
    return scg.get_rank_in_group("dp")


def param_sync(model, src_rank: int = 0, group=None, comm_group=None) -> None:
    """Broadcast parameters and buffers from ``src_rank`` (index inside the group); tensors tagged ``no_sync`` or tensor-parallel shards keep
    their local values.  ``comm_group`` is the reference's keyword for ``group``."""
    g = comm_group if comm_group is not None else (group if group is not None else scg.get_dp_group())
    if g is not None:
        C.broadcast_params(model, g, g.ranks[src_rank])


def grad_sync(params, group=None, scale=None, grad_avg: bool = True) -> None:
    """All-reduce gradients over the dp group, averaged unless ``grad_avg=False`` (or an explicit ``scale``).  Accepts a parameter list or
    the reference's optimizer ``param_groups``."""
    g = group if group is not None else scg.get_dp_group()
    params = list(params)
    if params and isinstance(params[0], dict):
        params = [q for grp in params for q in grp["params"] if not getattr(q, "tp_sharded", False)]
    if scale is None and not grad_avg:
        scale = 1.0
    C.fused_allreduce_gradients(params, g, scale=scale)


@torch.no_grad()
def all_reduce(tensor: torch.Tensor, op=dist.ReduceOp.SUM) -> torch.Tensor:
    """In-place all-reduce of a (metric) tensor over the dp group."""
    g = scg.get_dp_group()
    if g is not None and C.group_size(g) > 1 and g.process_group is not None:
        dist.all_reduce(tensor, op=op, group=g.process_group)
    return tensor

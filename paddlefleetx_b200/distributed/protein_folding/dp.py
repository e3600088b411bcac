"""Data-parallel helpers of the protein-folding stack (reference distributed/protein_folding/dp.py:41-107): broadcast
parameters from the dp source and all-reduce (+ scale) every gradient after backward."""
from __future__ import annotations

from ...parallel import comm_ops as C
from .scg import scg


def param_sync(model, src_rank: int = 0, group=None) -> None:
    g = group if group is not None else scg.get_dp_group()
    if g is not None:
        C.broadcast_params(model, g, g.ranks[src_rank])


def grad_sync(params, group=None, scale=None) -> None:
    g = group if group is not None else scg.get_dp_group()
    C.fused_allreduce_gradients(list(params), g, scale=scale)

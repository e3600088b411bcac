"""``wrap_with_fleet`` — pick the distributed wrappers for a (model, optimizer[, scaler]) triple.

Reference: eager_engine.py:274-323 / distributed/apis/strategy.py — ``fleet.distributed_model`` returns DataParallel /
TensorParallel / PipelineParallel / sharding wrappers and broadcasts initial parameters inside the mp / dp /
sharding groups.  Here: parameters are broadcast so replicas start identical, pipeline models get their 1F1B
scheduler, stage-3 models get the parameter-sharding wrapper; data-parallel / stage-1/2 gradient traffic is
owned by the flat optimizer, so there is no separate DataParallel reducer object.
"""
from __future__ import annotations

import torch

from ...parallel import comm_ops as C
from . import env


def sync_params_buffers(model: torch.nn.Module, comm_group, src_rank: int, skip_distributed: bool = False) -> None:
    C.broadcast_params(model, comm_group, src_rank, skip_distributed)


def broadcast_initial_parameters(model: torch.nn.Module, hcg) -> None:
    if env.world_size() == 1:
        return
    mp, dp, sh = hcg.get_model_parallel_group(), hcg.get_data_parallel_group(), hcg.get_sharding_parallel_group()
    if mp.nranks > 1:      # replicated (non-TP-sharded) tensors must agree inside the mp group
        C.broadcast_params(model, mp, hcg.get_model_parallel_group_src_rank(), skip_distributed=True)
    if sh.nranks > 1:
        C.broadcast_params(model, sh, hcg.get_sharding_parallel_group_src_rank())
    if dp.nranks > 1:
        C.broadcast_params(model, dp, hcg.get_data_parallel_group_src_rank())


def wrap_with_fleet(dist_config, model, optimizer=None, scaler=None):
    hcg = env.get_hcg()
    broadcast_initial_parameters(model, hcg)
    stage = dist_config.sharding.get("sharding_stage", 1)
    if dist_config.pp_degree > 1:
        from ...parallel.pipeline import PipelineParallel

        model = PipelineParallel(model, hcg, env.get_strategy())
    return model, optimizer, scaler

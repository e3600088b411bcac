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
    """ZeRO stage 2 / 3 -> ``wrap_sharding_2_3``; everything else -> ``wrap_3D_parallel`` (reference strategy.py:28-34)."""
    if dist_config.sharding.get("sharding_stage", 1) in (2, 3) and dist_config.sharding.get("sharding_degree", 1) > 1 and dist_config.pp_degree == 1:
        return wrap_sharding_2_3(dist_config, model, optimizer, scaler)
    return wrap_3D_parallel(dist_config, model, optimizer, scaler)


def wrap_sharding_2_3(dist_config, model, optimizer=None, scaler=None):
    """Group-sharded wrapping (reference strategy.py:37-72).  Replicas start identical; a stage-3 model gets the parameter-sharding wrapper
    (``parallel/sharding.py``: parameters live as shards, units are gathered around their use); stage 2 needs no model wrapper because the
    flat optimizer owns gradient reduce-scatter and the sharded update (it reads ``reduce_overlap`` / ``broadcast_overlap`` from the same
    ``Distributed.sharding`` section when ``build_optimizer`` constructs it).  Without an optimizer (evaluation) the model stays replicated."""
    hcg = env.get_hcg()
    assert dist_config.pp_degree == 1, "sharding stage2/3 will support pipeline parallel later"
    stage = dist_config.sharding.get("sharding_stage", 2)
    if not hasattr(model, "get_all_parameters"):             # not wrapped yet
        broadcast_initial_parameters(model, hcg)
        if stage == 3 and optimizer is not None and env.world_size() > 1 and hcg.get_sharding_parallel_world_size() > 1:
            from ...parallel.sharding import GroupShardedStage3

            model = GroupShardedStage3(model, hcg)
    return model, optimizer, scaler


def wrap_3D_parallel(dist_config, model, optimizer=None, scaler=None):
    """dp / mp / pp (and ZeRO-1) wrapping (reference strategy.py:75-94): initial parameter broadcast inside the mp / sharding / dp groups,
    the 1F1B scheduler around a pipeline model.  Data-parallel and stage-1 gradient traffic belongs to the flat optimizer, so there is no
    DataParallel reducer object and optimizer / scaler pass through."""
    hcg = env.get_hcg()
    inner = getattr(model, "_layers", None) if dist_config.pp_degree > 1 else None       # MixPrecisionLayer around a pipeline model
    target = inner if inner is not None else model
    broadcast_initial_parameters(target, hcg)
    if dist_config.pp_degree > 1:
        from ...parallel.pipeline import PipelineParallel

        model = PipelineParallel(target, hcg, env.get_strategy())
    return model, optimizer, scaler

"""Checkpoint save / load in the reference directory layout (eager_engine.py:717-830, distributed/apis/io.py:28-81):

    <output_dir>/epoch_{E}_step_{S}/[mp_{MM}_sharding_{SS}_pp_{PP}/]{model.pdparams, model_state.pdopt, meta_state.pdopt}

File names are kept ("same checkpoint layout"); payloads are ``torch.save`` pickles of plain dicts of CPU tensors.
Only ``dp_rank == 0`` writes.  ``meta_state.pdopt`` = {epoch, step, cuda_rng_state, rng_tracker, scaler}.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from ...parallel.rng import get_rng_state_tracker
from ...utils.log import logger
from . import env


def rank_subdir() -> str:
    if env.world_size() == 1:
        return ""
    h = env.get_hcg()
    return "mp_{:0>2d}_sharding_{:0>2d}_pp_{:0>2d}".format(h.get_model_parallel_rank(), h.get_sharding_parallel_rank(), h.get_stage_id())


def ckpt_dir(output_dir: str, epoch: int, step: int) -> str:
    return os.path.join(output_dir, f"epoch_{epoch}_step_{step}", rank_subdir())


def _layout() -> dict:
    if env.world_size() == 1:
        return {"mp": 1, "pp": 1, "sharding": 1, "dp": 1}
    h = env.get_hcg()
    return {"mp": h.get_model_parallel_world_size(), "pp": h.get_pipe_parallel_world_size(),
            "sharding": h.get_sharding_parallel_world_size(), "dp": h.get_data_parallel_world_size()}


def _cpu(obj):
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu()
    if isinstance(obj, dict):
        return {k: _cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_cpu(v) for v in obj)
    return obj


def save(output_dir: str, model: torch.nn.Module, optimizer=None, step: int = 0, epoch: int = 0, scaler=None, sharding_stage: int = 1) -> Optional[str]:
    if env.world_size() > 1 and env.get_hcg().get_data_parallel_rank() != 0:
        return None
    d = ckpt_dir(output_dir, epoch, step)
    os.makedirs(d, exist_ok=True)
    torch.save(_cpu(model.state_dict()), os.path.join(d, "model.pdparams"))
    if optimizer is not None:
        torch.save(_cpu(optimizer.state_dict()), os.path.join(d, "model_state.pdopt"))
    meta = {"epoch": epoch, "step": step, "cpu_rng_state": torch.get_rng_state(),
            "cuda_rng_state": torch.cuda.get_rng_state() if torch.cuda.is_available() else None,
            "rng_tracker": get_rng_state_tracker().get_states_tracker(),
            "scaler": scaler.state_dict() if scaler is not None else None,
            # tensor-parallel split axis of every sharded entry; lets utils/ckpt_convert.py merge / re-split offline
            "tp_axes": {n: int(getattr(p, "split_axis", 0)) for n, p in model.named_parameters() if getattr(p, "tp_sharded", False)},
            "layout": _layout()}
    torch.save(meta, os.path.join(d, "meta_state.pdopt"))
    logger.info(f"save model to {d}")
    return d


def load(ckpt_path: str, model: torch.nn.Module, optimizer=None, mode: str = "train", load_recovery: Optional[dict] = None, scaler=None) -> dict:
    d = os.path.join(ckpt_path, rank_subdir()) if rank_subdir() and not os.path.isfile(os.path.join(ckpt_path, "model.pdparams")) else ckpt_path
    mpath = os.path.join(d, "model.pdparams")
    if not os.path.isfile(mpath):
        raise ValueError(f"No model checkpoint file found in {d}.")
    state = torch.load(mpath, map_location="cpu", weights_only=False)
    own = model.state_dict()
    for k, v in own.items():
        if k not in state:
            raise KeyError(f"{k} is not found in the provided checkpoint")
        if state[k].dtype != v.dtype:
            state[k] = state[k].to(v.dtype)
    model.load_state_dict({k: state[k] for k in own}, strict=True)
    rec = load_recovery if load_recovery is not None else {}
    if mode == "train":
        opath, meta_path = os.path.join(d, "model_state.pdopt"), os.path.join(d, "meta_state.pdopt")
        if optimizer is not None:
            if not os.path.isfile(opath):
                raise ValueError(f"No optimizer checkpoint file found in {d}.")
            optimizer.set_state_dict(torch.load(opath, map_location="cpu", weights_only=False))
        if os.path.isfile(meta_path):
            meta = torch.load(meta_path, map_location="cpu", weights_only=False)
            rec.update(step=meta["step"], epoch=meta["epoch"], rng_state=meta.get("cuda_rng_state"), cpu_rng_state=meta.get("cpu_rng_state"),
                       rng_tracker=meta.get("rng_tracker"))
            if scaler is not None and meta.get("scaler") is not None:
                scaler.load_state_dict(meta["scaler"])
        else:
            raise ValueError(f"No meta checkpoint file found in {d}.")
    logger.info(f"successfully load checkpoints from {d}")
    return rec

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


def _expert_keys(model: torch.nn.Module) -> set:
    """State-dict keys of expert parameters.  Stacked expert parameters (moe/grouped_experts.py) appear in checkpoints under their per-expert
    names, so those are generated from the owning layer."""
    keys = {n for n, p in model.named_parameters() if getattr(p, "is_expert", False)}
    for name, mod in model.named_modules():
        ge = getattr(mod, "grouped", None)
        if ge is not None and hasattr(ge, "NAMES"):
            prefix = name + "." if name else ""
            for _, lin, attr in ge.NAMES:
                keys.update(f"{prefix}experts.{e}.{lin}.{attr}" for e in range(ge.num_expert))
    return keys


def _expert_replica_dir(d: str) -> Optional[str]:
    """Expert parallelism spreads the experts over the DATA-parallel ranks, so replicas 1.. hold weights (and optimizer moments) that replica 0
    does not: they get their own ``dp_XX`` sub-directory next to replica 0's files.  (The reference writes replica 0 only and cannot resume
    an expert-parallel run.)"""
    if env.world_size() == 1:
        return None
    r = env.get_hcg().get_data_parallel_rank()
    return os.path.join(d, f"dp_{r:02d}") if r != 0 else None


def _rng_states() -> dict:
    return {"cpu_rng_state": torch.get_rng_state(), "cuda_rng_state": torch.cuda.get_rng_state() if torch.cuda.is_available() else None,
            "rng_tracker": get_rng_state_tracker().get_states_tracker()}


def restore_rng(rec: dict) -> None:
    """Put the random-number streams recorded in a checkpoint (``load`` returns them in its recovery dict) back: device generator, host generator,
    the named tensor-parallel streams.  Call it right before the first step that trains."""
    if rec.get("rng_state") is not None and torch.cuda.is_available():
        torch.cuda.set_rng_state(rec["rng_state"])
    if rec.get("cpu_rng_state") is not None:
        torch.set_rng_state(rec["cpu_rng_state"])
    if rec.get("rng_tracker") is not None:
        get_rng_state_tracker().set_states_tracker(rec["rng_tracker"])


def save(output_dir: str, model: torch.nn.Module, optimizer=None, step: int = 0, epoch: int = 0, scaler=None, sharding_stage: int = 1) -> Optional[str]:
    d = ckpt_dir(output_dir, epoch, step)
    if env.world_size() > 1 and env.get_hcg().get_data_parallel_rank() != 0:
        # replicas 1.. only add what replica 0 cannot know: their random-number streams (dropout / routing noise differ per replica) and,
        # under expert parallelism, their experts with the matching optimizer state
        ed = _expert_replica_dir(d)
        os.makedirs(ed, exist_ok=True)
        torch.save(_rng_states(), os.path.join(ed, "meta_state.pdopt"))
        experts = _expert_keys(model)
        if experts:
            torch.save({k: v for k, v in _cpu(model.state_dict()).items() if k in experts}, os.path.join(ed, "model.pdparams"))
            if optimizer is not None:
                torch.save(_cpu(optimizer.state_dict()), os.path.join(ed, "model_state.pdopt"))
            logger.info(f"save this replica's experts to {ed}")
        return ed
    os.makedirs(d, exist_ok=True)
    torch.save(_cpu(model.state_dict()), os.path.join(d, "model.pdparams"))
    if optimizer is not None:
        torch.save(_cpu(optimizer.state_dict()), os.path.join(d, "model_state.pdopt"))
    meta = {"epoch": epoch, "step": step, **_rng_states(),
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
    ed = _expert_replica_dir(d)
    if ed is not None and os.path.isfile(os.path.join(ed, "model.pdparams")):          # this data-parallel replica's own experts
        state.update(torch.load(os.path.join(ed, "model.pdparams"), map_location="cpu", weights_only=False))
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
        if ed is not None and os.path.isfile(os.path.join(ed, "model_state.pdopt")):
            opath = os.path.join(ed, "model_state.pdopt")
        if optimizer is not None:
            if not os.path.isfile(opath):
                raise ValueError(f"No optimizer checkpoint file found in {d}.")
            optimizer.set_state_dict(torch.load(opath, map_location="cpu", weights_only=False))
        if os.path.isfile(meta_path):
            meta = torch.load(meta_path, map_location="cpu", weights_only=False)
            if ed is not None and os.path.isfile(os.path.join(ed, "meta_state.pdopt")):      # this replica's own random-number streams
                meta.update(torch.load(os.path.join(ed, "meta_state.pdopt"), map_location="cpu", weights_only=False))
            rec.update(step=meta["step"], epoch=meta["epoch"], rng_state=meta.get("cuda_rng_state"), cpu_rng_state=meta.get("cpu_rng_state"),
                       rng_tracker=meta.get("rng_tracker"))
            if scaler is not None and meta.get("scaler") is not None:
                scaler.load_state_dict(meta["scaler"])
        else:
            raise ValueError(f"No meta checkpoint file found in {d}.")
    logger.info(f"successfully load checkpoints from {d}")
    return rec


# ---------------------------------------------------------------------------------------------------- layout-annotated weights
def save_for_auto_inference(path_prefix: str, model: torch.nn.Module) -> Optional[str]:
    """Write this rank's weights with their distributed attributes: ``<prefix>_dist<rank>.pdparams`` (state dict of the local shards) and
    ``<prefix>_dist<rank>.pdattr`` (per tensor: process mesh ``[pp, mp]``, this rank's coordinates, ``dims_mapping`` = which tensor axis is split
    over the ``mp`` mesh axis, ``-1`` elsewhere).  The reference writes the same pair after every checkpoint so a training run on any hybrid
    layout can be served by the auto-parallel inference path on a different one (eager_engine.py:750-752 ``save_for_auto_inference``); here
    ``load_auto_inference`` re-assembles / re-splits the tensors for whatever layout the loading model has.  Replicas (dp / sharding rank > 0)
    write nothing."""
    world = env.world_size()
    if world > 1:
        h = env.get_hcg()
        if h.get_data_parallel_rank() != 0 or h.get_sharding_parallel_rank() != 0:
            return None
        mp, pp, mp_rank, pp_rank, rank = (h.get_model_parallel_world_size(), h.get_pipe_parallel_world_size(), h.get_model_parallel_rank(),
                                          h.get_stage_id(), env.global_rank())
    else:
        mp = pp = 1
        mp_rank = pp_rank = rank = 0
    os.makedirs(os.path.dirname(os.path.abspath(path_prefix)) or ".", exist_ok=True)
    sharded = {n: int(getattr(p, "split_axis", 0)) for n, p in model.named_parameters() if getattr(p, "tp_sharded", False)}
    state = _cpu(model.state_dict())
    attrs = {}
    for name, t in state.items():
        if not isinstance(t, torch.Tensor):
            continue
        mapping = [-1] * t.dim()
        if name in sharded and mp > 1:
            mapping[sharded[name]] = 1                      # mesh axis 1 = mp
        attrs[name] = {"process_shape": [pp, mp], "process_coord": [pp_rank, mp_rank], "dims_mapping": mapping}
    torch.save(state, f"{path_prefix}_dist{rank}.pdparams")
    torch.save({"mesh": [pp, mp], "coord": [pp_rank, mp_rank], "tensors": attrs}, f"{path_prefix}_dist{rank}.pdattr")
    logger.info(f"save auto-inference weights to {path_prefix}_dist{rank}.pdparams")
    return f"{path_prefix}_dist{rank}.pdparams"


def merge_auto_inference(path_prefix: str) -> dict:
    """All ``<prefix>_dist*.pdparams`` + ``.pdattr`` files -> one full (un-sharded) state dict: tensors split over ``mp`` are concatenated along
    their mapped axis in mesh-coordinate order, pipeline stages are unioned."""
    import glob

    files = sorted(glob.glob(f"{path_prefix}_dist*.pdattr"))
    if not files:
        raise FileNotFoundError(f"no {path_prefix}_dist*.pdattr files")
    pieces: dict = {}
    for f in files:
        attr = torch.load(f, map_location="cpu", weights_only=False)
        state = torch.load(f[:-len(".pdattr")] + ".pdparams", map_location="cpu", weights_only=False)
        for name, a in attr["tensors"].items():
            pieces.setdefault(name, []).append((a["process_coord"][1], a["dims_mapping"], state[name], a["process_coord"][0]))
    full = {}
    for name, parts in pieces.items():
        # tensor-parallel order first; among pipeline duplicates of a shared layer the earliest stage's copy wins (it is the trained one)
        parts.sort(key=lambda p: (p[0], p[3]))
        mapping = parts[0][1]
        axis = mapping.index(1) if 1 in mapping else None
        if axis is None:
            full[name] = parts[0][2]
        else:
            seen = {}
            for coord, _, t, _stage in parts:
                seen.setdefault(coord, t)             # the same shard may appear once per pipeline replica of a tied weight
            full[name] = torch.cat([seen[c] for c in sorted(seen)], dim=axis)
    return full


def load_auto_inference(path_prefix: str, model: torch.nn.Module) -> None:
    """Load layout-annotated weights into ``model`` whatever its own tensor-parallel degree: the full tensors are rebuilt, then every parameter
    the model marks ``tp_sharded`` takes this rank's slice along its ``split_axis``."""
    full = merge_auto_inference(path_prefix)
    if env.world_size() > 1:
        h = env.get_hcg()
        mp, mp_rank = h.get_model_parallel_world_size(), h.get_model_parallel_rank()
    else:
        mp, mp_rank = 1, 0
    axes = {n: int(getattr(p, "split_axis", 0)) for n, p in model.named_parameters() if getattr(p, "tp_sharded", False)}
    own = model.state_dict()
    out = {}
    for name, ref in own.items():
        if name not in full:
            raise KeyError(f"{name} is not found in {path_prefix}_dist*.pdparams")
        t = full[name]
        if name in axes and mp > 1:
            t = t.chunk(mp, dim=axes[name])[mp_rank]
        if tuple(t.shape) != tuple(ref.shape):
            raise ValueError(f"{name}: stored {tuple(t.shape)} does not fit {tuple(ref.shape)} at mp={mp}")
        out[name] = t.to(ref.dtype)
    model.load_state_dict(out, strict=True)
    logger.info(f"loaded auto-inference weights from {path_prefix}_dist*.pdparams (mp={mp})")


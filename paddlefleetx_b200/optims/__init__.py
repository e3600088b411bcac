"""Builders: ``build_lr_scheduler`` / ``build_grad_clip`` / ``build_optimizer`` from the ``Optimizer:`` YAML block.

Registry semantics follow ppfleetx/optims/__init__.py:29-74 (name looked up in this namespace, remaining keys
passed as kwargs), without ``eval``.
"""
from __future__ import annotations

import copy
from typing import Optional

import torch

from ..utils.log import logger
from . import grad_clip as _gc
from . import lr_scheduler as _lr
from . import optimizer as _opt
from .grad_clip import ClipGradByGlobalNorm, ClipGradForMOEByGlobalNorm
from .lr_scheduler import (CosineAnnealingWithWarmupDecay, CosineDecay, LinearDecayWithWarmup, LRScheduler, MultiStepDecay,
                           ViTLRScheduler)
from .optimizer import Adam, AdamW, FusedAdamW, Momentum

__all__ = ["build_lr_scheduler", "build_grad_clip", "build_optimizer", "FusedAdamW", "AdamW", "Adam", "Momentum",
           "ClipGradByGlobalNorm", "ClipGradForMOEByGlobalNorm", "LRScheduler"]


def _lookup(module, name: str):
    if not hasattr(module, name):
        raise ValueError(f"{name} is not defined in {module.__name__}")
    return getattr(module, name)


def build_lr_scheduler(lr_config):
    if lr_config is None:
        return None
    if isinstance(lr_config, (int, float)):
        return float(lr_config)
    cfg = copy.deepcopy(dict(lr_config))
    if "name" not in cfg:
        return float(cfg.get("learning_rate", cfg.get("max_lr", 1e-3)))
    name = cfg.pop("name")
    sched = _lookup(_lr, name)(**cfg)
    logger.debug(f"build lr ({name}) success..")
    return sched


def build_grad_clip(grad_clip_config, hcg=None):
    if grad_clip_config is None:
        return None
    cfg = copy.deepcopy(dict(grad_clip_config))
    name = cfg.pop("name", "ClipGradByGlobalNorm")
    return _lookup(_gc, name)(hcg=hcg, **cfg)


def build_optimizer(config, model: torch.nn.Module, lr_scheduler=None, hcg=None, dist_config=None, amp_config=None):
    cfg = copy.deepcopy(dict(config))
    cfg.pop("lr", None)
    grad_clip = build_grad_clip(cfg.pop("grad_clip", None), hcg)
    name = cfg.pop("name", "FusedAdamW")
    cfg.pop("tensor_fusion", None)      # the flat layout is unconditional here
    extra = {}
    if dist_config is not None:
        sh = dist_config.get("sharding", {})
        extra.update(sharding_stage=sh.get("sharding_stage", 1), reduce_overlap=sh.get("reduce_overlap", False),
                     broadcast_overlap=sh.get("broadcast_overlap", False), use_p2p=sh.get("use_p2p", False),
                     bucket_mb=sh.get("bucket_mb", 512), offload=bool(sh.get("sharding_offload", False)) or bool(cfg.pop("offload", False)))
    extra["step_overlap"] = bool(cfg.pop("step_overlap", False))
    if amp_config is not None:
        extra["use_main_grad"] = bool(amp_config.get("use_main_grad", False))
    if hasattr(model, "optimizer_named_parameters"):            # ZeRO-3 wrapper: optimise the shards
        named = model.optimizer_named_parameters()
        extra.update(params_are_shards=True, apply_decay_param_fun=model.shard_decay_fn())
    else:
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    cls = _lookup(_opt, name)
    opt = cls(learning_rate=lr_scheduler if lr_scheduler is not None else cfg.pop("learning_rate", 1e-3),
              named_parameters=named, grad_clip=grad_clip, hcg=hcg, **{**cfg, **extra})
    logger.debug(f"build optimizer ({name}) success..")
    return opt

"""``build_dataloader(cfg.Data, mode)`` — dataset / sampler / collate by name from the YAML block
(reference ppfleetx/data/__init__.py:28-119), over ``torch.utils.data.DataLoader`` with pinned memory."""
from __future__ import annotations

import copy
import random
from typing import Optional

import numpy as np
import torch

from ..distributed.apis import env
from ..utils.log import logger
from . import dataset as _datasets
from .sampler import batch_sampler as _samplers
from .utils import batch_collate_fn as _collate      # task collate functions + the Stack / Pad / Tuple / Dict primitives


def _lookup(mod, name):
    if not hasattr(mod, name):
        raise ValueError(f"{name} is not defined in {mod.__name__}")
    return getattr(mod, name)


def build_dataset(config, mode: str):
    cfg = copy.deepcopy(dict(config[mode]["dataset"] if "dataset" in config[mode] else config[mode]))
    name = cfg.pop("name")
    ds = _lookup(_datasets, name)(**cfg)
    logger.debug(f"build dataset({name}) success...")
    return ds


def build_batch_sampler(config, dataset):
    if config is None:
        return None
    cfg = copy.deepcopy(dict(config))
    name = cfg.pop("name")
    return _lookup(_samplers, name)(dataset, **cfg)


def build_collate_fn(spec):
    if spec is None:
        return None
    if isinstance(spec, str):
        return _lookup(_collate, spec)
    cfg = copy.deepcopy(dict(spec))
    name = cfg.pop("name")
    return _lookup(_collate, name)(**cfg)


def _worker_init(worker_id: int) -> None:
    base = env.get_dp_seed() or 0
    np.random.seed(base + worker_id)
    random.seed(base + worker_id)


def build_dataloader(config, mode: str):
    assert mode in ("Train", "Eval", "Test"), "Dataset mode should be Train, Eval, Test"
    if mode not in config:
        return None
    dataset = build_dataset(config, mode)
    sampler = build_batch_sampler(config[mode].get("sampler"), dataset)
    lcfg = copy.deepcopy(dict(config[mode].get("loader", {})))
    collate = build_collate_fn(lcfg.pop("collate_fn", None))
    lcfg.pop("return_list", None)
    num_workers = int(lcfg.pop("num_workers", 0))
    pin = bool(lcfg.pop("pin_memory", torch.cuda.is_available()))
    kwargs = dict(num_workers=num_workers, collate_fn=collate, pin_memory=pin, worker_init_fn=_worker_init if num_workers else None,
                  persistent_workers=bool(num_workers) and bool(lcfg.pop("persistent_workers", False)))
    if num_workers:
        kwargs["prefetch_factor"] = int(lcfg.pop("prefetch_factor", 2))
    if sampler is not None:
        loader = torch.utils.data.DataLoader(dataset, batch_sampler=sampler, **kwargs)
    else:
        loader = torch.utils.data.DataLoader(dataset, batch_size=int(lcfg.pop("batch_size", 1)), shuffle=bool(lcfg.pop("shuffle", False)),
                                             drop_last=bool(lcfg.pop("drop_last", False)), **kwargs)
    logger.debug(f"build dataloader({mode}) success...")
    return loader


def build_auto_dataset(config, mode: str):
    return build_dataset(config, mode)

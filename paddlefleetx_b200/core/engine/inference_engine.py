"""``InferenceEngine`` — serve an exported model (reference core/engine/inference_engine.py:41-271).

``InferenceEngine(model_dir, mp_degree, tensorrt_config=None).predict(list | dict of numpy arrays)`` loads
``model_dir/rank_{r}/model.pdmodel|.pdiparams`` (``r`` = this process's model-parallel rank; ``auto_dist{r}`` prefix accepted),
rebuilds the module from the recipe and runs it on the GPU.  Distributed inference (mp_degree > 1) uses the normal
``torch.distributed`` groups of the training runtime instead of Paddle-Inference's ring-id CSV.  ``TensorRTConfig`` is kept
as a documented no-op (one backend: our own kernels — fp16/bf16/int8 precision is a property of the exported artifact).
"""
from __future__ import annotations

import os
from collections.abc import Mapping, Sequence
from typing import List, Optional

import numpy as np
import torch

from ...utils.config import AttrDict, _wrap
from ...utils.export import load_recipe
from ...utils.log import logger


class TensorRTConfig:
    """Accepted for API parity (max_batch_size, workspace_size, min_subgraph_size, precision, use_static, use_calib_mode,
    collect_shape, shape_range_info_filename); precision 'int8' selects the int8 tcgen05 path if the artifact is quantised."""

    def __init__(self, max_batch_size=1, workspace_size=1 << 30, min_subgraph_size=3, precision="fp16", use_static=False, use_calib_mode=False,
                 collect_shape=False, shape_range_info_filename=None, **unused):
        self.max_batch_size, self.workspace_size, self.min_subgraph_size = max_batch_size, workspace_size, min_subgraph_size
        self.precision, self.use_static, self.use_calib_mode = precision, use_static, use_calib_mode
        self.collect_shape, self.shape_range_info_filename = collect_shape, shape_range_info_filename


class InferenceEngine:
    def __init__(self, model_dir: str, mp_degree: int = 1, tensorrt_config: Optional[TensorRTConfig] = None, device: Optional[str] = None):
        self.model_dir, self.mp_degree, self.tensorrt_config = model_dir, mp_degree, tensorrt_config
        self.rank = 0
        if mp_degree > 1:
            import torch.distributed as dist

            from ...distributed.apis import env

            env.init_process_group("gpu")
            self.rank = env.get_hcg().get_model_parallel_rank() if dist.is_initialized() else 0
        self._check_model()
        self.device = torch.device(device) if device else (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
        self._init_predictor()

    def _check_model(self):
        cands = [os.path.join(self.model_dir, f"rank_{self.rank}"), self.model_dir]
        for d in cands:
            if os.path.isdir(d) and any(f.endswith(".pdmodel") for f in os.listdir(d)):
                self.rank_dir = d
                return
        raise ValueError(f"no exported model found under {self.model_dir} (looked for rank_{self.rank}/*.pdmodel)")

    def _init_predictor(self):
        from ...models import _REGISTRY
        import importlib

        recipe, params = load_recipe(self.rank_dir)
        self.recipe = recipe
        cfg = AttrDict({k: _wrap(v) for k, v in recipe.items() if isinstance(v, dict)})
        cfg.setdefault("Global", AttrDict(device="gpu" if self.device.type == "cuda" else "cpu", seed=1024, global_batch_size=1, local_batch_size=1, micro_batch_size=1))
        cfg.Global["device"] = "gpu" if self.device.type == "cuda" else "cpu"
        cfg.setdefault("Engine", AttrDict())
        cfg.Engine.setdefault("mix_precision", AttrDict(enable=False))
        cfg.Engine.setdefault("max_steps", 1); cfg.Engine.setdefault("eval_freq", -1); cfg.Engine.setdefault("eval_iters", 0); cfg.Engine.setdefault("test_iters", 0)
        cfg.setdefault("Distributed", AttrDict(dp_degree=1, mp_degree=self.mp_degree, pp_degree=1, sharding=AttrDict(sharding_degree=1, sharding_stage=1)))
        cfg.setdefault("Optimizer", AttrDict(lr=AttrDict()))
        name = cfg.Model.module
        base_cls = getattr(importlib.import_module(_REGISTRY[name]), name)
        # the recipe already holds post-processed values: skip the post-processing in a throw-away subclass (patching the class itself would
        # switch it off for every later build_module of the same task in this process)
        module_cls = type(name, (base_cls,), {"process_configs": lambda self_, c: c})
        module = module_cls(cfg)
        model = module.model
        prune = ((recipe.get("Compress") or {}).get("Prune") or {})
        if prune.get("enable", False):              # a structurally pruned export: rebuild the pruned shapes before the weights are loaded
            from ...utils.compression_helper import prune_model

            prune_model(model, prune)
        if recipe.get("quant"):
            from ...utils.compression_helper import convert_to_int8, quant_model

            quant_model(model, {})
            convert_to_int8(model)
        if recipe.get("smooth_quant"):
            from ...utils.smoothquant import install_empty

            install_empty(model, recipe["smooth_quant"]["layers"], dtype=next(model.parameters()).dtype)
        state = torch.load(params, map_location="cpu", weights_only=False)
        own = model.state_dict()
        model.load_state_dict({k: v.to(own[k].dtype) if k in own and hasattr(v, "to") else v for k, v in state.items() if k in own}, strict=False)
        self.module, self.model = module, model.to(self.device).eval()
        self.input_names = [s.get("name", f"input_{i}") for i, s in enumerate(recipe.get("input_spec") or [])]
        logger.info(f"InferenceEngine: loaded {name} from {self.rank_dir} on {self.device}")

    def input_names_list(self) -> List[str]:
        return list(self.input_names)

    @torch.no_grad()
    def predict(self, data):
        if isinstance(data, Mapping):
            arrays = [data[n] for n in self.input_names] if all(n in data for n in self.input_names) else list(data.values())
        elif isinstance(data, Sequence):
            arrays = list(data)
        else:
            arrays = [data]
        tensors = [torch.as_tensor(np.asarray(a)).to(self.device) for a in arrays]
        out = self.model(*tensors)
        outs = out if isinstance(out, (tuple, list)) else [out]
        return {f"output_{i}": o.detach().float().cpu().numpy() if o.is_floating_point() else o.detach().cpu().numpy() for i, o in enumerate(outs)}

"""Inference export: a self-contained per-rank artifact ``<dir>/{model.pdmodel, model.pdiparams}``.

The reference traces the dygraph model to a static program (``paddle.jit.to_static`` + ``paddle.jit.save``,
ppfleetx/utils/export.py:24-72).  There is no tracing compiler on the B200 path: ``model.pdmodel`` is a JSON *recipe* (how to
rebuild the module from the registry: module name, YAML sections, input spec, quantisation info) and ``model.pdiparams`` the
weights; ``InferenceEngine`` rebuilds the eager module and runs it (CUDA-graph captured decode loop for generation).
"""
from __future__ import annotations

import json
import os
from typing import Optional

import torch

from .log import logger

FORMAT = "pfx-b200-export-v1"


def _plain(obj):
    if isinstance(obj, dict):
        return {str(k): _plain(v) for k, v in obj.items() if not str(k).startswith("_")}
    if isinstance(obj, (list, tuple)):
        return [_plain(v) for v in obj]
    if isinstance(obj, (str, int, float, bool)) or obj is None:
        return obj
    return str(obj)


def export_inference_model(model: torch.nn.Module, input_spec, save_dir: str, save_prefix: str = "model", configs=None, quant: bool = False,
                           smooth_quant: Optional[dict] = None) -> str:
    os.makedirs(save_dir, exist_ok=True)
    recipe = {"format": FORMAT, "input_spec": _plain(input_spec), "quant": bool(quant)}
    if smooth_quant:
        recipe["smooth_quant"] = _plain(smooth_quant)      # alpha / shift / the rewritten layers: InferenceEngine rebuilds the same structure
    if configs is not None:
        for sec in ("Model", "Generation", "Distributed", "Global", "Inference", "Data", "Offline_Eval", "Compress"):
            if sec in configs:
                recipe[sec] = _plain(configs[sec])
        recipe["Engine"] = {"mix_precision": _plain(configs.Engine.mix_precision)}
    if quant:
        from .compression_helper import convert_to_int8

        model = convert_to_int8(model)
    with open(os.path.join(save_dir, save_prefix + ".pdmodel"), "w") as f:
        json.dump(recipe, f, indent=1)
    state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    torch.save(state, os.path.join(save_dir, save_prefix + ".pdiparams"))
    logger.info(f"exported inference model to {save_dir} ({len(state)} tensors)")
    return save_dir


def load_recipe(model_dir: str, prefix: Optional[str] = None):
    names = [f for f in os.listdir(model_dir) if f.endswith(".pdmodel")]
    if len(names) != 1:
        raise ValueError(f"expected exactly one .pdmodel in {model_dir}, found {names}")
    prefix = prefix or names[0][:-len(".pdmodel")]
    with open(os.path.join(model_dir, prefix + ".pdmodel")) as f:
        recipe = json.load(f)
    assert recipe.get("format") == FORMAT, f"unknown export format in {model_dir}"
    params = os.path.join(model_dir, prefix + ".pdiparams")
    if not os.path.isfile(params):
        raise ValueError(f"missing {params}")
    return recipe, params

"""Module registry: ``build_module(cfg)`` instantiates ``cfg.Model.module`` (reference models/__init__.py:30-34 uses
``eval``; here an explicit name table)."""
from __future__ import annotations

import importlib

_REGISTRY = {
    "GPTModule": "paddlefleetx_b200.models.language_model.language_module",
    "GPTFinetuneModule": "paddlefleetx_b200.models.language_model.finetune_module",
    "GPTGenerationModule": "paddlefleetx_b200.models.language_model.generation_module",
    "GPTEvalModule": "paddlefleetx_b200.models.language_model.eval_module",
    "MoEModule": "paddlefleetx_b200.models.language_model.moe_module",
    "GPTModuleAuto": "paddlefleetx_b200.models.language_model.auto_module",
    "GPTGenerationModuleAuto": "paddlefleetx_b200.models.language_model.auto_module",
    "ErnieModule": "paddlefleetx_b200.models.language_model.ernie.ernie_module",
    "ErnieSeqClsModule": "paddlefleetx_b200.models.language_model.ernie.ernie_module",
    "ErnieModuleAuto": "paddlefleetx_b200.models.language_model.ernie.ernie_module",
    "ErnieSeqClsModuleAuto": "paddlefleetx_b200.models.language_model.ernie.ernie_module",
    "GeneralClsModule": "paddlefleetx_b200.models.vision_model.general_classification_module",
    "GeneralClsModuleAuto": "paddlefleetx_b200.models.vision_model.general_classification_module",
    "MOCOModule": "paddlefleetx_b200.models.vision_model.moco_module",
    "MOCOClsModule": "paddlefleetx_b200.models.vision_model.moco_module",
    "ImagenModule": "paddlefleetx_b200.models.multimodal_model.multimodal_module",
}


def register_module(name: str, module_path: str) -> None:
    _REGISTRY[name] = module_path


def build_module(config):
    name = config.Model.module
    if name not in _REGISTRY:
        raise ValueError(f"unknown module {name}; known: {sorted(_REGISTRY)}")
    cls = getattr(importlib.import_module(_REGISTRY[name]), name)
    return cls(config)

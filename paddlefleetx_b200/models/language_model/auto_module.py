"""``GPTModuleAuto`` / ``GPTGenerationModuleAuto`` (reference gpt/auto/auto_module.py:26-145 builds a static-graph model whose weights carry
``auto.shard_tensor(weight, mesh, [None, 'mp'])`` annotations and lets Paddle's planner complete the program).

Here the sharding lives in the layers, so the "auto" modules are the hybrid modules plus what the annotations stood for:

* the process mesh recorded by ``get_auto_config`` (``Distributed.mesh`` = [pp, dp x sharding, mp]) must match the topology the layers
  were built on — a mismatch means the YAML's mesh and degrees disagree, which the reference would only notice as a wrong placement;
* ``shard_spec()`` reports, per parameter, the mesh axis it is split along (the information ``auto.shard_tensor`` carried), which
  ``tools/auto_export.py`` writes next to an exported model and tests use to check that every tensor-parallel weight is annotated;
* a planner-chosen layout (``Distributed.auto_layout``) is logged with its predicted step time and memory.
"""
from __future__ import annotations

from typing import Dict, List, Optional

from ...utils.log import logger
from .generation_module import GPTGenerationModule
from .language_module import GPTModule


class _AutoMixin:
    def _check_mesh(self, configs) -> None:
        d = configs.get("Distributed", {}) or {}
        mesh = d.get("mesh")
        if mesh is not None:
            want = [int(d.get("pp_degree", 1)), int(d.get("dp_degree", 1)) * int(d.get("sharding", {}).get("sharding_degree", 1)), int(d.get("mp_degree", 1))]
            if list(mesh.shape) != want:
                raise ValueError(f"Distributed.mesh {list(mesh.shape)} does not match the degrees (pp, dp x sharding, mp) = {want}")
        plan = d.get("plan")
        if plan is not None:
            logger.info(f"[auto] planner layout: {plan.describe}")

    def shard_spec(self) -> Dict[str, List[Optional[str]]]:
        """``{parameter name: [mesh axis or None per tensor dim]}`` — 'mp' on the split axis of tensor-parallel weights, replicated otherwise."""
        spec = {}
        for name, p in self.model.named_parameters():
            dims: List[Optional[str]] = [None] * p.dim()
            if getattr(p, "tp_sharded", False) and p.dim() > 0:
                dims[int(getattr(p, "split_axis", 0))] = "mp"
            spec[name] = dims
        return spec


class GPTModuleAuto(_AutoMixin, GPTModule):
    def __init__(self, configs):
        self._check_mesh(configs)
        super().__init__(configs)


class GPTGenerationModuleAuto(_AutoMixin, GPTGenerationModule):
    def __init__(self, configs):
        self._check_mesh(configs)
        super().__init__(configs)

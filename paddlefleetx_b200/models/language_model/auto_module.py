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


class ProcessMesh:
    """An n-dimensional arrangement of ranks with named axes — the object ``auto.shard_tensor`` annotations refer to in the reference.
    ``mesh[i]`` is the sub-mesh of stage ``i`` along the first axis; ``processes`` the flat rank list."""

    def __init__(self, mesh, dim_names: Optional[List[str]] = None):
        import numpy as np

        self._array = np.asarray(mesh)
        self.dim_names = list(dim_names) if dim_names is not None else [f"d{i}" for i in range(self._array.ndim)]
        assert len(self.dim_names) == self._array.ndim, f"{len(self.dim_names)} names for a {self._array.ndim}-d mesh"

    @property
    def shape(self) -> List[int]:
        return list(self._array.shape)

    @property
    def processes(self) -> List[int]:
        return [int(r) for r in self._array.reshape(-1)]

    @property
    def ndim(self) -> int:
        return self._array.ndim

    def get_dim_size(self, name: str) -> int:
        return self._array.shape[self.dim_names.index(name)]

    def __getitem__(self, idx) -> "ProcessMesh":
        sub = self._array[idx]
        return ProcessMesh(sub if sub.ndim else sub.reshape(1), self.dim_names[1:] if sub.ndim else ["d0"])

    def __eq__(self, other):
        return isinstance(other, ProcessMesh) and self.dim_names == other.dim_names and self._array.shape == other._array.shape \
            and bool((self._array == other._array).all())

    def __repr__(self):
        return f"ProcessMesh(shape={self.shape}, dim_names={self.dim_names})"


def process_mesh_config(config):
    """The mesh object the reference's auto models receive as ``Model.mesh`` (auto_utils.py:24-108): axes with degree 1 are dropped, the
    order is ``pp``, ``dp``, ``mp``; ``mesh[stage]`` is that pipeline stage's (dp, mp) sub-mesh (or the whole mesh without pipeline),
    ``mesh.stages(num_layers)`` maps layers to stages, ``mesh.dp`` / ``mesh.mp`` name the axes that exist (``None`` otherwise)."""
    import numpy as np

    class Mesh:
        def __init__(self, config):
            self.config = config
            axes = [(n, int(config.get(f"{n}_degree", 1) or 1)) for n in ("pp", "dp", "mp")]
            live = [(n, d) for n, d in axes if d > 1]
            count = int(np.prod([d for _, d in live])) if live else 1
            ranks = np.arange(count)
            if live:
                self.process_mesh = ProcessMesh(ranks.reshape([d for _, d in live]), [n for n, _ in live])
            else:
                self.process_mesh = ProcessMesh(ranks)
            names = [n for n, _ in live]
            self.dp_dim = "dp" if "dp" in names else None
            self.mp_dim = "mp" if "mp" in names else None

        def __getitem__(self, idx):
            return self.process_mesh[idx] if "pp" in self.process_mesh.dim_names else self.process_mesh

        def stages(self, num_layers: int) -> List[int]:
            per_stage = num_layers // int(self.config.get("pp_degree", 1) or 1)
            return [i // per_stage for i in range(num_layers)]

        @property
        def dp(self):
            return self.dp_dim

        @property
        def mp(self):
            return self.mp_dim

    return Mesh(config)


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


class LanguageModuleAuto(_AutoMixin, GPTModule):
    """Base of the auto modules (reference gpt/auto/auto_module.py:26-60): the hybrid module plus the mesh check and ``shard_spec``."""

    def __init__(self, configs):
        self._check_mesh(configs)
        super().__init__(configs)


class GPTModuleAuto(_AutoMixin, GPTModule):
    def __init__(self, configs):
        self._check_mesh(configs)
        super().__init__(configs)


class GPTGenerationModuleAuto(_AutoMixin, GPTGenerationModule):
    def __init__(self, configs):
        self._check_mesh(configs)
        super().__init__(configs)

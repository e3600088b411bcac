"""``AutoEngine`` (reference core/engine/auto_engine.py:39-209 wraps Paddle's static-graph semi-auto-parallel engine).

There is no tracing planner here: an "auto" config describes a process mesh (pp, dp, mp) that maps one-to-one onto the
hybrid topology, so ``AutoEngine`` is the eager engine with the mesh validated — ``tools/auto.py`` / ``tools/auto_export.py``
keep working with the same YAML files.  ``tune`` reports the candidate layouts ranked by an analytic cost model."""
from __future__ import annotations

from .eager_engine import EagerEngine


class AutoEngine(EagerEngine):
    def __init__(self, configs, module=None, mode="train"):
        mesh = configs.Distributed.get("mesh")
        if mesh is not None:
            d = configs.Distributed
            assert list(mesh.shape) == [d.pp_degree, d.dp_degree * d.sharding.sharding_degree, d.mp_degree], "mesh / degree mismatch"
        super().__init__(configs, module, mode=mode)

    def tune(self, tune_data_loader=None):
        from ...utils.layout_planner import rank_layouts

        return rank_layouts(self._configs)

    def export_from_prog(self):
        return self.export()

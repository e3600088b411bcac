"""``AutoEngine`` (reference core/engine/auto_engine.py:39-209 wraps Paddle's static-graph semi-auto-parallel engine and its
``OptimizationTuner``).

An "auto" config describes a process mesh (pp, dp, mp) that maps one-to-one onto the hybrid topology, so ``AutoEngine`` is the eager engine
with the mesh validated — ``tools/auto.py`` / ``tools/auto_export.py`` keep working with the same YAML files.  What the reference's planner
decides by propagating ``shard_tensor`` annotations is decided here by a cost model over the concrete parallel implementations:
``Distributed.auto_layout: True`` makes ``utils.config.get_auto_config`` ask ``utils/layout_planner.py`` for the fastest layout that fits
180 GB per GPU (degrees, ZeRO stage, micro-batch, recompute) and derive the config from it; the chosen plan is recorded under
``Distributed.plan``.  ``tune`` covers the two things the reference's ``Tuning`` section drives:

* ``Tuning.tuning_recompute`` / ``Tuning.tuning_micro_batch`` — a *measured* search: every recompute setting (off / ``core_attn`` /
  ``full_attn`` / ``full``) and / or every micro-batch size that divides the local batch is built on the configured mesh, steps ``[profile_start_step, profile_end_step]`` are timed (device-synchronised, max over ranks) with the peak allocator
  footprint recorded, and the fastest candidate under the memory limit is written back into the config;
* otherwise — the planner's ranked table of every feasible (dp, sharding + stage, mp, pp) layout with its predicted step time, memory and
  time breakdown (compute / TP collectives / pipeline bubble / exposed DP traffic / exposed optimizer).
"""
from __future__ import annotations

import copy
import gc
import time
from typing import Dict, Iterable, List, Optional

import torch

from ...utils.log import logger
from .eager_engine import EagerEngine

RECOMPUTE_CANDIDATES = ({"use_recompute": False}, {"use_recompute": True, "recompute_granularity": "core_attn"},
                        {"use_recompute": True, "recompute_granularity": "full_attn"}, {"use_recompute": True, "recompute_granularity": "full"})


class AutoEngine(EagerEngine):
    def __init__(self, configs, module=None, mode="train"):
        mesh = configs.Distributed.get("mesh")
        if mesh is not None:
            d = configs.Distributed
            assert list(mesh.shape) == [d.pp_degree, d.dp_degree * d.sharding.sharding_degree, d.mp_degree], "mesh / degree mismatch"
        super().__init__(configs, module, mode=mode)

    # ------------------------------------------------------------------ tuning
    def tune(self, tune_data_loader: Optional[Iterable] = None) -> List[Dict]:
        tuning = self._configs.get("Tuning", {}) or {}
        if tuning.get("tuning_recompute", False) or tuning.get("tuning_micro_batch", False):
            assert tune_data_loader is not None, "Tuning.tuning_recompute / tuning_micro_batch measure real steps: pass the training data loader"
            return self._tune_measured(tune_data_loader, int(tuning.get("profile_start_step", 1)), int(tuning.get("profile_end_step", 5)),
                                       tuning.get("memory_limit_gb"), recompute=bool(tuning.get("tuning_recompute", False)),
                                       micro_batch=bool(tuning.get("tuning_micro_batch", False)))
        from ...utils.layout_planner import rank_layouts

        return rank_layouts(self._configs)

    def _profile_candidate(self, overrides: Dict, batches: List, start: int, end: int) -> Dict:
        from ...distributed.apis import env
        from ...models import build_module

        cfg = copy.deepcopy(self._configs)
        overrides = dict(overrides)
        micro = overrides.pop("micro_batch_size", None)
        cfg.Model.update(overrides)
        if micro is not None:                     # the engine accumulates local_batch / micro_batch micro-steps per optimizer step
            cfg.Global.micro_batch_size = int(micro)
            cfg.Engine.accumulate_steps = int(cfg.Global.local_batch_size) // int(micro)
            overrides["micro_batch_size"] = int(micro)
        cfg.Engine.save_load.update({"save_steps": -1, "ckpt_dir": None})
        cuda = torch.cuda.is_available() and str(cfg.Global.get("device", "gpu")) != "cpu"
        row = dict(overrides, status="ok")
        engine = loss = None
        if cuda:                                  # whatever is resident already (this engine's own model, allocator leftovers) is not the candidate's
            gc.collect()
            torch.cuda.empty_cache()
            torch.cuda.synchronize()
            torch.cuda.reset_peak_memory_stats()
            resident = torch.cuda.memory_allocated()
        try:
            env.set_seed(cfg.Global.seed)
            engine = EagerEngine(configs=cfg, module=build_module(cfg))
            t0 = None
            for i, batch in enumerate(batches[: end + 1]):
                if i == start:
                    if cuda:
                        torch.cuda.synchronize()
                    t0 = time.perf_counter()
                loss = engine.train_step(batch)
            if cuda:
                torch.cuda.synchronize()
            dt = torch.tensor([(time.perf_counter() - t0) / (end - start + 1)], dtype=torch.float64)
            mem = torch.tensor([(torch.cuda.max_memory_allocated() - resident) / 2 ** 30 if cuda else 0.0], dtype=torch.float64)
            if env.world_size() > 1:              # a candidate is as slow / as large as its worst rank
                import torch.distributed as dist

                dev = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
                dt, mem = dt.to(dev), mem.to(dev)
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                dist.all_reduce(mem, op=dist.ReduceOp.MAX)
            row.update(step_s=float(dt), peak_mem_gb=float(mem), final_loss=float(loss))
        except torch.cuda.OutOfMemoryError:
            row.update(status="oom", step_s=float("inf"), peak_mem_gb=float("inf"))
        engine = loss = None                      # drop the candidate (model, flat optimizer buffers, hooks) before the next one is built
        gc.collect()
        if cuda:
            torch.cuda.empty_cache()
        return row

    def _micro_batch_candidates(self) -> List[int]:
        lb = int(self._configs.Global.local_batch_size)
        pp = int(self._configs.Distributed.get("pp_degree", 1) or 1)
        return [m for m in range(1, lb + 1) if lb % m == 0 and (pp == 1 or lb // m >= pp)]

    def _tune_measured(self, loader: Iterable, start: int, end: int, memory_limit_gb: Optional[float], recompute: bool = True,
                       micro_batch: bool = False) -> List[Dict]:
        """Build and time every candidate (recompute setting x micro-batch size, whichever the Tuning section asks for) on the configured
        mesh; the fastest one under the memory limit is written back into the config."""
        assert 0 <= start <= end, "need 0 <= profile_start_step <= profile_end_step"
        batches = []
        for batch in loader:
            batches.append(batch)
            if len(batches) > end:
                break
        assert len(batches) > end, f"the loader yields {len(batches)} batches, profiling needs {end + 1}"
        if memory_limit_gb is None and torch.cuda.is_available():
            memory_limit_gb = 0.92 * torch.cuda.get_device_properties(torch.cuda.current_device()).total_memory / 2 ** 30
        rc = [dict(c) for c in RECOMPUTE_CANDIDATES] if recompute else [{}]
        mb = [{"micro_batch_size": m} for m in self._micro_batch_candidates()] if micro_batch else [{}]
        rows = [self._profile_candidate({**a, **b}, batches, start, end) for a in rc for b in mb]
        for r in rows:
            r["fits"] = r["status"] == "ok" and (memory_limit_gb is None or r["peak_mem_gb"] <= memory_limit_gb)
        rows.sort(key=lambda r: (not r["fits"], r["step_s"]))
        best = rows[0]
        if best["fits"]:
            self._configs.Model.update({k: best[k] for k in ("use_recompute", "recompute_granularity") if k in best})
            if "micro_batch_size" in best:
                self._configs.Global.micro_batch_size = best["micro_batch_size"]
                self._configs.Engine.accumulate_steps = int(self._configs.Global.local_batch_size) // best["micro_batch_size"]
            logger.info(f"[tune] selected " + ", ".join(f"{k}={best[k]}" for k in ("use_recompute", "recompute_granularity", "micro_batch_size") if k in best)
                        + f" ({best['step_s'] * 1e3:.1f} ms/step, peak {best['peak_mem_gb']:.1f} GB)")
        else:
            logger.warning("[tune] no candidate fits the memory limit; configuration left unchanged")
        return rows

    def _tune_recompute(self, loader: Iterable, start: int, end: int, memory_limit_gb: Optional[float]) -> List[Dict]:
        return self._tune_measured(loader, start, end, memory_limit_gb, recompute=True, micro_batch=False)

    def export_from_prog(self):
        return self.export()

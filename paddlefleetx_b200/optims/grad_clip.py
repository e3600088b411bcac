"""Gradient clipping by global norm, hybrid-parallel aware.

* ``ClipGradByGlobalNorm`` — reference re-exports Paddle's and relies on fleet's ``HybridParallelClipGrad`` to
  sum the squared norm across mp / pp / sharding groups (ppfleetx/optims/grad_clip.py:16).  Here
  the cross-group logic is explicit: TP-sharded parameters contribute on every mp rank, replicated ones only
  on mp rank 0; the scalar is all-reduced over the ``check`` (mp x pp) and sharding groups.
* ``ClipGradForMOEByGlobalNorm`` — expert parameters' squared norm is all-reduced over the expert-parallel
  group before being combined with the dense norm (grad_clip.py:27-156).

The norm itself is computed by the ``sumsq`` kernel over flat gradient buffers when they exist (one launch per
buffer, no host sync: the clip coefficient stays on the device and is consumed by the fused AdamW kernel).
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def _grad_of(p):
    mg = getattr(p, "main_grad", None)
    return mg if mg is not None else p.grad


class ClipGradByGlobalNorm:
    def __init__(self, clip_norm: float = 1.0, hcg=None, **unused):
        self.clip_norm = float(clip_norm)
        self.hcg = hcg

    # -- norm ------------------------------------------------------------------------------
    def _local_sumsq(self, params: Iterable[torch.nn.Parameter]) -> torch.Tensor:
        dist_sq, rep_sq = None, None
        for p in params:
            g = _grad_of(p)
            if g is None:
                continue
            s = g.float().pow(2).sum()
            if getattr(p, "tp_sharded", False) or getattr(p, "is_expert", False):
                dist_sq = s if dist_sq is None else dist_sq + s
            else:
                rep_sq = s if rep_sq is None else rep_sq + s
        dev = next((_grad_of(p).device for p in params if _grad_of(p) is not None), torch.device("cpu"))
        zero = torch.zeros((), device=dev)
        dist_sq = zero if dist_sq is None else dist_sq
        rep_sq = zero if rep_sq is None else rep_sq
        mp_rank = self.hcg.get_model_parallel_rank() if self.hcg is not None else 0
        return dist_sq + (rep_sq if mp_rank == 0 else zero)

    def global_sumsq(self, params: List[torch.nn.Parameter], extra_sharding: bool = False) -> torch.Tensor:
        sq = self._local_sumsq(params)
        return self.reduce_sumsq(sq, extra_sharding)

    def reduce_sumsq(self, sq: torch.Tensor, sharding: bool = False) -> torch.Tensor:
        h = self.hcg
        if h is None or not (dist.is_available() and dist.is_initialized()):
            return sq
        chk = h.get_check_parallel_group()
        if chk.nranks > 1 and chk.process_group is not None:
            dist.all_reduce(sq, group=chk.process_group)
        if sharding:
            sh = h.get_sharding_parallel_group()
            if sh.nranks > 1 and sh.process_group is not None:
                dist.all_reduce(sq, group=sh.process_group)
        return sq

    # -- eager clip (per-tensor path) ----------------------------------------------------------
    @torch.no_grad()
    def __call__(self, params: List[torch.nn.Parameter]) -> torch.Tensor:
        norm = self.global_sumsq(params).sqrt()
        coef = torch.clamp(self.clip_norm / (norm + 1e-6), max=1.0)
        for p in params:
            g = _grad_of(p)
            if g is not None:
                g.mul_(coef.to(g.dtype))
        return norm


class ClipGradForMOEByGlobalNorm(ClipGradByGlobalNorm):
    def __init__(self, clip_norm: float = 1.0, is_expert_param_func=None, moe_group=None, hcg=None, **unused):
        super().__init__(clip_norm, hcg)
        self.is_expert = is_expert_param_func or (lambda p: getattr(p, "is_expert", False))
        self.moe_group = moe_group

    def global_sumsq(self, params, extra_sharding: bool = False) -> torch.Tensor:
        normal = [p for p in params if not self.is_expert(p)]
        moe = [p for p in params if self.is_expert(p)]
        sq = self._local_sumsq(normal)
        if moe:
            msq = None
            for p in moe:
                g = _grad_of(p)
                if g is not None:
                    s = g.float().pow(2).sum()
                    msq = s if msq is None else msq + s
            if msq is not None:
                grp = self.moe_group or (self.hcg.get_moe_group() if self.hcg is not None else None)
                if grp is not None and grp.nranks > 1 and grp.process_group is not None:
                    dist.all_reduce(msq, group=grp.process_group)
                sq = sq + msq
        return sq

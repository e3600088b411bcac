"""Analytic layout planner used by ``AutoEngine.tune``: enumerate (dp, mp, pp, sharding) factorisations of the world and
rank them by estimated step time = compute / (peak * efficiency) + exposed TP / PP / DP communication."""
from __future__ import annotations

from ..parallel.topology import all_axis_products
from .config import world_size_hint


def rank_layouts(cfg, peak_flops: float = 1.45e15, link_bw: float = 7.0e11):
    m = cfg.Model
    L, h, s = m.get("num_layers", 12), m.get("hidden_size", 768), cfg.Data.Train.dataset.get("max_seq_len", 1024) if "Data" in cfg else 1024
    params = 12 * L * h * h
    world = world_size_hint()
    b = cfg.Global.micro_batch_size
    out = []
    for dp, mp, pp, sd in all_axis_products(world):
        if L % pp or m.get("num_attention_heads", 12) % mp:
            continue
        tokens = b * s
        compute = 6 * params * tokens / (mp * pp) / peak_flops
        tp_comm = 0 if mp == 1 else 4 * (L / pp) * tokens * h * 2 * (mp - 1) / mp / link_bw
        bubble = (pp - 1) / max(cfg.Engine.accumulate_steps, 1)
        dp_comm = 2 * params * 2 / (mp * pp) * (dp * sd - 1) / max(dp * sd, 1) / link_bw / max(cfg.Engine.accumulate_steps, 1)
        mem = params * (2 + 2 + 12 / max(sd, 1)) / (mp * pp)
        out.append(dict(dp=dp, mp=mp, pp=pp, sharding=sd, est_step_s=(compute + tp_comm) * (1 + bubble) + 0.3 * dp_comm, est_mem_gb=mem / 2 ** 30))
    return sorted([o for o in out if o["est_mem_gb"] < 170], key=lambda o: o["est_step_s"])

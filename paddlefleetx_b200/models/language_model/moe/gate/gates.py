"""MoE gates: ``NaiveGate`` (linear + top-k), ``GShardGate`` (top-2, load-balance loss, capacity, random second
expert), ``SwitchGate`` (top-1, multiplicative jitter, capacity) — reference moe/gate/*.py."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .....parallel import comm_ops as C
from ..utils import limit_by_capacity, number_count, random_routing


class BaseGate(nn.Module):
    def __init__(self, num_expert: int, group=None):
        super().__init__()
        self.group = group
        self.world_size = C.group_size(group)
        self.num_expert = num_expert
        self.tot_expert = num_expert * self.world_size
        self.loss = None

    def set_loss(self, loss):
        self.loss = loss

    def get_loss(self, clear: bool = True):
        loss = self.loss
        if clear:
            self.loss = None
        return loss

    @property
    def has_loss(self) -> bool:
        return self.loss is not None


class NaiveGate(BaseGate):
    def __init__(self, d_model: int, num_expert: int, group=None, topk: int = 2, dtype=None, device=None):
        super().__init__(num_expert, group)
        self.gate = nn.Linear(d_model, self.tot_expert, dtype=dtype, device=device)
        for p in self.gate.parameters():
            p.is_gate = True               # tag read by ClipGradForMOEByGlobalNorm / param broadcast (reference renames params)
        self.top_k = topk

    def forward(self, inp, return_all_scores: bool = False):
        score = self.gate(inp)
        val, idx = torch.topk(score, k=self.top_k, dim=-1, largest=True, sorted=True)
        if return_all_scores:
            return val, idx, score
        return val, idx


class GShardGate(NaiveGate):
    def __init__(self, d_model: int, num_expert: int, topk: int = 2, capacity=(1.2, 2.4), random_routing: bool = True, group=None,
                 dtype=None, device=None):
        assert topk == 2, "topk should be 2 in gshard"
        super().__init__(d_model, num_expert, group, topk, dtype, device)
        self.capacity = capacity
        self.random_routing = random_routing

    def forward(self, x):
        topk_val, topk_idx, score = super().forward(x, return_all_scores=True)
        s = score.shape[0]
        # the per-expert slot counts serve both the load-balance loss (fraction of slots routed to each expert) and the capacity limit:
        # counted once (every op in here is a kernel launch, 24 layers deep: the 8-GPU MoE step is bound by launches, not by math)
        lec = number_count(topk_idx, self.tot_expert)
        c_e = lec.float() / s
        m_e = F.softmax(score.float(), dim=1).mean(0)
        self.set_loss((c_e * m_e).mean() * (self.num_expert ** 2))
        cap = math.ceil(self.capacity[0 if self.training else 1] * x.shape[0])
        _, _, topk_idx = limit_by_capacity(topk_idx, self.num_expert, self.world_size, cap, group=self.group, lec=lec)
        if self.random_routing:
            prob = torch.rand(s, dtype=torch.float32, device=x.device)
            topk_idx = random_routing(topk_idx, topk_val, prob)
        return topk_val, topk_idx


class SwitchGate(NaiveGate):
    def __init__(self, d_model: int, num_expert: int, topk: int = 1, switch_eps: float = 0.1, capacity=(1.2, 2.4), group=None, dtype=None,
                 device=None):
        assert topk == 1, "topk should be 1 in switch"
        super().__init__(d_model, num_expert, group, 1, dtype, device)
        self.switch_eps = switch_eps
        self.capacity = capacity

    def forward(self, inp):
        score = self.gate(inp)
        if self.training:
            noise = torch.rand_like(score) * 2 * self.switch_eps + 1.0 - self.switch_eps
            score = score + noise
        score = F.softmax(score.float(), dim=-1)
        top1_score, top1_idx = torch.topk(score, k=1, dim=-1, largest=True)
        cap = math.ceil(self.capacity[0 if self.training else 1] * inp.shape[0])
        _, _, top1_idx = limit_by_capacity(top1_idx, self.num_expert, self.world_size, cap, group=self.group)
        valid = (top1_idx >= 0).sum().clamp(min=1).float()
        frac_expert = torch.zeros(self.tot_expert, dtype=torch.float32, device=inp.device)
        ok = top1_idx.reshape(-1) >= 0
        frac_expert.index_add_(0, top1_idx.reshape(-1).clamp(min=0), ok.float())                     # dropped slots add 0 (no host sync)
        frac_expert = frac_expert / valid
        prob_expert = score.sum(0) / valid
        self.set_loss((frac_expert * prob_expert).sum() * self.tot_expert)
        return top1_score.to(inp.dtype), top1_idx

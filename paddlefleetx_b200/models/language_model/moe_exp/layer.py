"""``MoE`` — user-facing capacity-based MoE block (reference moe_exp/layer.py:30-107): TopKGate + Experts + MOELayer, optional
residual MoE (dense MLP mixed with the expert output through a learned 2-way coefficient)."""
from __future__ import annotations

import copy
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .experts import Experts
from .sharded_moe import MOELayer, TopKGate


class MoE(nn.Module):
    def __init__(self, hidden_size: int, expert: nn.Module, num_experts: int = 1, ep_size: int = 1, k: int = 1, capacity_factor: float = 1.0,
                 eval_capacity_factor: float = 1.0, min_capacity: int = 4, use_residual: bool = False, noisy_gate_policy: Optional[str] = None,
                 drop_tokens: bool = True, use_rts: bool = False, enable_expert_tensor_parallelism: bool = False, ep_group=None):
        super().__init__()
        assert num_experts % ep_size == 0, f"Number of experts ({num_experts}) should be divisible by expert parallel size ({ep_size})"
        assert noisy_gate_policy in (None, "None", "Jitter", "RSample"), "Unsupported noisy_gate_policy: " + str(noisy_gate_policy)
        self.use_residual = use_residual
        self.enable_expert_tensor_parallelism = enable_expert_tensor_parallelism
        self.ep_size, self.num_experts = ep_size, num_experts
        self.num_local_experts = num_experts // ep_size
        self.expert_group_name = f"ep_size_{ep_size}"
        experts = Experts(expert, self.num_local_experts, self.expert_group_name)
        self.gate = TopKGate(hidden_size, num_experts, k, capacity_factor, eval_capacity_factor, min_capacity,
                             None if noisy_gate_policy == "None" else noisy_gate_policy, drop_tokens, use_rts)
        self.fleetx_moe = MOELayer(self.gate, experts, self.expert_group_name, ep_size, self.num_local_experts)
        if ep_group is not None:
            self.set_expert_parallel_group(ep_group)
        if use_residual:
            self.mlp = copy.deepcopy(expert)
            self.coefficient = nn.Linear(hidden_size, 2)

    def set_expert_parallel_group(self, group) -> None:
        assert group.nranks == self.ep_size, f"ep group has {group.nranks} ranks, layer was built for ep_size={self.ep_size}"
        self.fleetx_moe._set_ep_group(group)

    def forward(self, hidden_states: torch.Tensor, used_token: Optional[torch.Tensor] = None):
        """returns (output, l_aux, exp_counts)"""
        out = self.fleetx_moe(hidden_states, used_token)
        if self.use_residual:
            mlp = self.mlp(hidden_states)
            mlp = mlp[0] if isinstance(mlp, tuple) else mlp
            coef = F.softmax(self.coefficient(hidden_states), dim=-1)
            out = out * coef[..., 0:1] + mlp * coef[..., 1:]
        return out, self.fleetx_moe.l_aux, self.fleetx_moe.exp_counts

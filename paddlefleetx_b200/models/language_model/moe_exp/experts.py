"""Local expert bank (reference moe_exp/experts.py:26-55): ``num_local_experts`` deep copies of a prototype expert; input
``[ep, E_local, C, M]`` is chunked along the local-expert axis, each chunk goes through its expert, outputs are re-stacked."""
import copy

import torch
import torch.nn as nn


class Experts(nn.Module):
    def __init__(self, expert: nn.Module, num_local_experts: int = 1, expert_group_name=None):
        super().__init__()
        self.experts = nn.ModuleList([copy.deepcopy(expert) for _ in range(num_local_experts)])
        self.num_local_experts = num_local_experts
        for e in self.experts:
            for p in e.parameters():                 # expert weights are private to their rank: no data-parallel sync
                p.is_expert = True
                p.no_sync = True
                p.group_name = expert_group_name

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        chunks = inputs.chunk(self.num_local_experts, dim=1)
        outs = []
        for chunk, expert in zip(chunks, self.experts):
            out = expert(chunk)
            outs.append(out[0] if isinstance(out, tuple) else out)
        return torch.cat(outs, dim=1)

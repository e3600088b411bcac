"""``MoEModule`` — GPT with expert-parallel MoE FFNs (reference language_module.py:736-830).

Differences from ``GPTModule``: the expert-parallel world is the fused dp x mp ``moe`` group (pp and sharding must be
1, comm_groups.py:133-137); the loss adds ``balance_loss_weight * mean(gate aux losses)``; non-expert parameters are
broadcast inside mp/dp at start while expert parameters stay rank-private.
"""
from __future__ import annotations

import torch

from ...distributed.apis import env
from ...utils.log import logger
from .language_module import GPTModule
from .moe.gate import BaseGate


class MoEModule(GPTModule):
    def __init__(self, configs):
        d = configs.Distributed
        assert d.pp_degree == 1 and d.sharding.sharding_degree == 1, "MoE runs with pp_degree == 1 and sharding_degree == 1"
        configs.Model.setdefault("moe_configs", {})
        configs.Model.moe_configs.setdefault("expert_mode", True)
        super().__init__(configs)
        self.balance_loss_weight = float(configs.Engine.get("balance_loss_weight", configs.Model.moe_configs.get("balance_loss_weight", 1.0)))
        self.gates = [m for m in self.model.modules() if isinstance(m, BaseGate)]
        n_exp = sum(p.numel() for p in self.model.parameters() if getattr(p, "is_expert", False))
        logger.info(f"MoE: {len(self.gates)} gated layers, {n_exp / 1e6:.1f} M expert parameters on this rank")

    def training_step(self, batch):
        tokens, position_ids, labels, loss_mask = batch
        preds = self(tokens, position_ids)
        loss = self.loss_fn(preds, labels, loss_mask)
        aux = [g.get_loss() for g in self.gates if g.has_loss]
        if aux and self.balance_loss_weight:
            loss = loss + self.balance_loss_weight * torch.stack([a.float() for a in aux]).mean()
        return loss

    def validation_step(self, batch):
        out = super().validation_step(batch)
        for g in self.gates:
            g.get_loss()
        return out

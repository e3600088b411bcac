"""GPT-MoE pre-training pieces (reference examples/transformer/models/GPT/pretrain_moe/impls.py): same loop as dense GPT, plus
the gate balance loss and the expert / non-expert parameter split (experts are not synchronised across data-parallel ranks)."""
import torch

from paddlefleetx_b200.models.language_model.moe_module import MoEModule


def build_module(config):
    return MoEModule(config)


def fit_impl(config, batch, module, optimizer):
    module.model.train()
    loss = module.training_step(batch)          # LM loss + balance_loss_weight * mean(gate aux losses)
    loss.backward()
    optimizer.step()
    optimizer.clear_grad()
    return loss.detach()


@torch.no_grad()
def eval_impl(config, batch, module):
    module.model.eval()
    return module.validation_step(batch)

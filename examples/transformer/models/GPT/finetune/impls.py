"""GPT sequence-classification fine-tuning pieces (reference examples/transformer/models/GPT/finetune/impls.py): the model,
loss and metric come from the task module's component builders; the loop in run.py owns stepping and logging."""
import torch

from paddlefleetx_b200.models.language_model.finetune_module import GPTFinetuneModule


def build_components(config):
    module = GPTFinetuneModule(config)          # builds model (+ pretrained load, qkv layout conversion), loss and metrics
    return module.model, module.loss_fn, module.metric, module


def fit_impl(config, batch, module, optimizer):
    module.model.train()
    loss = module.training_step(batch)
    loss.backward()
    optimizer.step()
    optimizer.clear_grad()
    return loss.detach()


@torch.no_grad()
def eval_impl(config, batch, module):
    module.model.eval()
    return module.validation_step(batch)

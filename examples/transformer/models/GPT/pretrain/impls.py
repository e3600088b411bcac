"""Pieces of the explicit GPT pre-training loop (reference examples/transformer/models/GPT/pretrain/impls.py): model
construction for the topology, one fit step with gradient accumulation, evaluation, parameter-group broadcast."""
import torch

from paddlefleetx_b200.distributed.apis import amp as amp_api
from paddlefleetx_b200.distributed.apis import env
from paddlefleetx_b200.models.language_model.gpt import model as gpt
from paddlefleetx_b200.parallel.tp_layers import allreduce_sequence_parallel_grads


def build_model(config):
    from paddlefleetx_b200.models.language_model.language_module import _device, _param_dtype, model_kwargs

    d = config.Distributed
    nranks = env.world_size()
    vocab = gpt.vocab_size_with_padding(config.Model.get("vocab_size", 50304), config.Model.get("vocab_size_divisible_unit", 128), d.mp_degree)
    config.Model["vocab_size"] = vocab
    kw = model_kwargs(config)
    dtype, device = _param_dtype(config), _device(config)
    mp_group = env.get_hcg().get_model_parallel_group() if (nranks > 1 and d.mp_degree > 1) else None
    if nranks > 1 and d.pp_degree > 1:
        from paddlefleetx_b200.models.language_model.gpt.pipe import GPTForPretrainingPipe

        model = GPTForPretrainingPipe(env.get_hcg(), mp_group, dtype=dtype, device=device, **kw)
        return model, None
    model = gpt.GPTForPretraining(gpt.GPTModel(mp_group=mp_group, dtype=dtype, device=device, **kw))
    return model, gpt.GPTPretrainingCriterion(mp_group)


def _split(batch, n):
    return [[t.chunk(n, 0)[i] for t in batch] for i in range(n)]


def model_forward_backward(config, batch, model, loss_fn, optimizer, scaler=None):
    """Micro-batched forward/backward; gradient reduction is deferred to the last micro-batch (``optimizer.no_sync``)."""
    acc = config.Engine.accumulate_steps
    mp = config.Engine.mix_precision
    total = 0.0
    micro = _split(batch, acc) if acc > 1 else [batch]
    for i, mb in enumerate(micro):
        tokens, position_ids, labels, loss_mask = mb
        ctx = optimizer.no_sync() if (i < acc - 1 and hasattr(optimizer, "no_sync")) else torch.enable_grad()
        with ctx:
            with amp_api.autocast_context(mp.get("enable", False), mp.get("dtype", "bfloat16"), mp.get("level", "O2"), tokens.device.type):
                loss = loss_fn(model(tokens, position_ids), labels, loss_mask)
            lb = scaler.scale(loss) if scaler is not None else loss
            (lb / acc).backward()
        total = total + loss.detach()
    if config.Distributed.mp_degree > 1 and config.Model.get("sequence_parallel", False):
        allreduce_sequence_parallel_grads(model)
    return total / acc


def fit_impl(config, batch, model, loss_fn, optimizer, scaler=None):
    model.train()
    if config.Distributed.pp_degree > 1:
        tokens, position_ids, labels, loss_mask = batch
        model._prepare_training(batch, optimizer, None)
        loss = model.forward_backward_pipeline([(tokens, position_ids), (labels, loss_mask)], scaler)
    else:
        loss = model_forward_backward(config, batch, model, loss_fn, optimizer, scaler)
    if scaler is not None:
        scaler.step(optimizer)
        scaler.update()
    else:
        optimizer.step()
    optimizer.clear_grad()
    return loss


@torch.no_grad()
def eval_impl(config, batch, model, loss_fn):
    model.eval()
    tokens, position_ids, labels, loss_mask = batch
    mp = config.Engine.mix_precision
    with amp_api.autocast_context(mp.get("enable", False), mp.get("dtype", "bfloat16"), mp.get("level", "O2"), tokens.device.type):
        if config.Distributed.pp_degree > 1:
            return model.eval_batch([(tokens, position_ids), (labels, loss_mask)], compute_loss=True)
        return loss_fn(model(tokens, position_ids), labels, loss_mask)

"""Offline evaluation without the engine (reference examples/transformer/models/GPT/offline-eval/impls.py:36-260): model construction, the
per-batch scoring step, and the final report.  Datasets (``LM_Eval_Dataset`` — overlapping WikiText windows —, ``Lambada_Eval_Dataset``) and
the detokeniser are the framework's (``paddlefleetx_b200/data/dataset/gpt_dataset.py``)."""
import math

import torch
import torch.nn.functional as F

from paddlefleetx_b200.data.dataset.gpt_dataset import Lambada_Eval_Dataset, LM_Eval_Dataset  # noqa: F401
from paddlefleetx_b200.models.language_model.gpt import model as gpt


def build_model(config):
    from paddlefleetx_b200.models.language_model.language_module import _device, _param_dtype, model_kwargs

    kw = model_kwargs(config)
    kw["use_flash_attn"] = False          # the evaluation passes an explicit attention mask
    model = gpt.GPTForPretraining(gpt.GPTModel(dtype=_param_dtype(config), device=_device(config), **kw))
    return model.eval()


@torch.no_grad()
def eval_impl(config, batch, model):
    """-> (score, info): summed masked cross-entropy for perplexity, number of fully-correct samples for the cloze task."""
    tokens, loss_mask, attention_mask, position_ids, labels, info = batch
    logits = model(tokens, position_ids, (1.0 - attention_mask.float()) * -1e4).float()
    if config.Offline_Eval.get("cloze_eval", False):
        hit = ((logits.argmax(-1) == labels) | (loss_mask == 0)).all(-1)
        return hit.float().sum(), info
    ce = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1), reduction="none").view_as(labels)
    return (ce * loss_mask).sum(), info


def report(config, total_score, info):
    ev = config.Offline_Eval
    if ev.get("cloze_eval", False):
        n = int(info[0][0])
        return "validation results on {} | number correct: {:.4E} | total examples: {:.4E} | avg accuracy: {:.4E}".format(
            ev.get("eval_path"), total_score, n, total_score / max(n, 1))
    n_orig, n_tok = int(info[0][0]), int(info[0][1])
    avg = total_score / max(n_tok - 1, 1)
    ratio = (n_tok - 1) / max(n_orig - 1, 1)
    return "validation results on {} | avg loss: {:.4E} | ppl: {:.4E} | adjusted ppl: {:.4E} | token ratio: {} |".format(
        ev.get("eval_path"), avg, math.exp(min(20, avg)), math.exp(min(20, avg * ratio)), ratio)

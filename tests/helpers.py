import os

from paddlefleetx_b200.utils import config as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = os.path.join(ROOT, "paddlefleetx_b200", "configs", "nlp", "gpt", "pretrain_gpt_small_synthetic.yaml")

TINY = ["Global.device=cpu", "Model.num_layers=2", "Model.hidden_size=64", "Model.num_attention_heads=4", "Model.ffn_hidden_size=128",
        "Model.vocab_size=512", "Model.vocab_size_divisible_unit=8", "Data.Train.dataset.max_seq_len=32", "Data.Eval.dataset.max_seq_len=32",
        "Model.max_position_embeddings=32", "Model.hidden_dropout_prob=0.0", "Model.attention_probs_dropout_prob=0.0",
        "Engine.mix_precision.enable=False", "Optimizer.lr.max_lr=1e-2", "Optimizer.lr.min_lr=1e-3", "Optimizer.lr.warmup_rate=0.0",
        "Engine.logging_freq=1"]


def tiny_gpt_config(overrides=(), nranks=None):
    return C.get_config(SMALL, TINY + list(overrides), show=False, nranks=nranks)


def build_engine(cfg):
    from paddlefleetx_b200.core import EagerEngine
    from paddlefleetx_b200.distributed.apis import env
    from paddlefleetx_b200.models import build_module

    env.init_dist_env(cfg)
    env.set_seed(cfg.Global.seed)
    module = build_module(cfg)
    return EagerEngine(configs=cfg, module=module)


def synthetic_batches(cfg, n, rank=0, seed=0):
    import torch

    g = torch.Generator().manual_seed(seed)
    b, s, v = cfg.Global.global_batch_size, cfg.Data.Train.dataset.max_seq_len, cfg.Model.vocab_size
    out = []
    for _ in range(n):
        toks = torch.randint(0, v, (b, s + 1), generator=g)
        pos = torch.arange(s).unsqueeze(0).expand(b, s).contiguous()
        out.append([toks[:, :-1].contiguous(), pos, toks[:, 1:].contiguous(), torch.ones(b, s)])
    return out

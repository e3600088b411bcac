"""WikiText perplexity / LAMBADA accuracy with an explicit loop (reference examples/transformer/models/GPT/offline-eval/run.py).

    python examples/transformer/models/GPT/offline-eval/run.py -c examples/transformer/models/GPT/offline-eval/configs/eval_gpt_345M_single_card.yaml \
        -o Engine.save_load.ckpt_dir=./ckpt/345M -o Offline_Eval.eval_path=./wikitext-103/wiki.valid.tokens -o Offline_Eval.cloze_eval=False
"""
import os
import sys
import time

__dir__ = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(__dir__, "..", "..", "..", "..", "..")))
sys.path.insert(0, os.path.abspath(os.path.join(__dir__, "..", "..", "..")))

import torch  # noqa: E402

import impls  # noqa: E402
from paddlefleetx_b200.data import build_dataloader  # noqa: E402
from paddlefleetx_b200.distributed.apis import env, io  # noqa: E402
from paddlefleetx_b200.utils.log import logger  # noqa: E402
from utils import config as cfg_utils  # noqa: E402


def main(argv=None):
    args = cfg_utils.parse_args(argv)
    config = cfg_utils._get_config(args.config, overrides=args.override, show=False)      # the evaluation recipe needs no pre-training post-processing
    ev = config.Offline_Eval
    ds = config.Data.Eval.dataset
    ds.update(name="Lambada_Eval_Dataset" if ev.get("cloze_eval", False) else "LM_Eval_Dataset", input_dir=ev.get("eval_path"),
              max_seq_len=ev.get("max_seq_len", 1024))
    if not ev.get("cloze_eval", False):
        ds["overlapping_eval"] = ev.get("overlapping_eval", 32)
    config.Data.Eval.loader["batch_size"] = ev.get("batch_size", 8)
    env.set_seed(config.Global.seed)
    model = impls.build_model(config)
    ckpt = config.Engine.save_load.get("ckpt_dir")
    if ckpt:
        io.load(ckpt, model, None, "eval")
    loader = build_dataloader(config.Data, "Eval")
    device = next(model.parameters()).device
    total, info, t0 = 0.0, None, time.time()
    for step, batch in enumerate(loader):
        batch = [b.to(device) if torch.is_tensor(b) else b for b in batch]
        score, info = impls.eval_impl(config, batch, model)
        total += float(score)
        if (step + 1) % ev.get("logging_freq", 10) == 0:
            logger.info("[eval] batch: %d, %s: %.9f, speed: %.2f step/s" % (step + 1, "number correct" if ev.get("cloze_eval", False) else "loss",
                                                                           total, (step + 1) / (time.time() - t0)))
    line = impls.report(config, total, info)
    logger.info(line)
    return line


if __name__ == "__main__":
    main()

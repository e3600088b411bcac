"""GPT-MoE pre-training with an explicit loop (reference .../pretrain_moe/run.py); expert-parallel world = dp group."""
import os
import sys

__dir__ = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(__dir__, "..", "..", "..", "..", "..")))
sys.path.insert(0, os.path.abspath(os.path.join(__dir__, "..", "..", "..")))

import time  # noqa: E402

import torch  # noqa: E402

import impls  # noqa: E402
from paddlefleetx_b200.distributed.apis import env, io, strategy  # noqa: E402
from paddlefleetx_b200.utils.log import logger  # noqa: E402
from utils import components as cpn  # noqa: E402
from utils import config as cfg_utils  # noqa: E402


def main():
    args = cfg_utils.parse_args()
    config = cfg_utils.get_config(args.config, overrides=args.override)
    if env.world_size() > 1:
        env.init_dist_env(config)
    env.set_seed(config.Global.seed)
    module = impls.build_module(config)
    model = module.model
    lr = cpn.build_lr_scheduler(config.Optimizer.lr)
    optimizer = cpn.build_optimizer(config.Optimizer, model, lr, dist_config=config.Distributed, amp_config=config.Engine.mix_precision)
    if env.world_size() > 1:
        model, optimizer, _ = strategy.wrap_with_fleet(config.Distributed, model, optimizer, None)
    device = next(model.parameters()).device
    ckpt = config.Engine.save_load.get("ckpt_dir")
    rec = io.load(ckpt, model, optimizer, "train") if ckpt else None     # each dp replica reads its own experts / optimizer slice / RNG
    start_step = rec.get("step", 0) if rec else 0
    loader = cpn.build_dataloader(config.Data, "Train")
    t0 = time.time()
    for step, batch in enumerate(loader):
        if step < start_step:
            continue
        if step >= config.Engine.max_steps:
            break
        if rec is not None:           # the gate's noise and dropout continue from the checkpointed streams
            io.restore_rng(rec)
            rec = None
        loss = impls.fit_impl(config, [t.to(device) for t in batch], module, optimizer)
        lr.step()
        if (step + 1) % config.Engine.logging_freq == 0:
            dt = (time.time() - t0) / config.Engine.logging_freq
            logger.train("[train] step: %d/%d, loss: %.9f, avg_batch_cost: %.5f sec" % (step + 1, config.Engine.max_steps, float(loss), dt))
            t0 = time.time()
        if config.Engine.save_load.save_steps > 0 and (step + 1) % config.Engine.save_load.save_steps == 0:
            io.save(config.Engine.save_load.output_dir, model, optimizer, step=step + 1, epoch=0)


if __name__ == "__main__":
    main()

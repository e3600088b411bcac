"""GLUE fine-tuning with an explicit loop (reference examples/transformer/models/GPT/finetune/run.py)."""
import os
import sys

__dir__ = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(__dir__, "..", "..", "..", "..", "..")))
sys.path.insert(0, os.path.abspath(os.path.join(__dir__, "..", "..", "..")))

import torch  # noqa: E402

import impls  # noqa: E402
from paddlefleetx_b200.distributed.apis import env, io  # noqa: E402
from paddlefleetx_b200.utils.log import logger  # noqa: E402
from utils import components as cpn  # noqa: E402
from utils import config as cfg_utils  # noqa: E402


def main():
    args = cfg_utils.parse_args()
    config = cfg_utils.get_config(args.config, overrides=args.override)
    if env.world_size() > 1:
        env.init_dist_env(config)
    env.set_seed(config.Global.seed)
    model, loss_fn, metric, module = impls.build_components(config)
    train_loader = cpn.build_dataloader(config.Data, "Train")
    valid_loader = cpn.build_dataloader(config.Data, "Eval")
    steps_per_epoch = len(train_loader)
    config.Optimizer.lr.setdefault("decay_steps", steps_per_epoch * config.Engine.num_train_epochs)
    lr = cpn.build_lr_scheduler(config.Optimizer.lr)
    optimizer = cpn.build_optimizer(config.Optimizer, model, lr, dist_config=config.Distributed, amp_config=config.Engine.mix_precision)
    device = next(model.parameters()).device
    for epoch in range(config.Engine.num_train_epochs):
        for step, batch in enumerate(train_loader):
            batch = [t.to(device) if isinstance(t, torch.Tensor) else t for t in batch]
            loss = impls.fit_impl(config, batch, module, optimizer)
            lr.step()
            if (step + 1) % config.Engine.logging_freq == 0:
                logger.train("[train] epoch: %d, step: %d/%d, loss: %.6f, lr: %.3e" % (epoch, step + 1, steps_per_epoch, float(loss), optimizer.get_lr()))
        for batch in valid_loader:
            impls.eval_impl(config, [t.to(device) if isinstance(t, torch.Tensor) else t for t in batch], module)
        module.validation_epoch_end({"epoch": epoch})
        io.save(config.Engine.save_load.output_dir, model, optimizer, step=steps_per_epoch, epoch=epoch)


if __name__ == "__main__":
    main()

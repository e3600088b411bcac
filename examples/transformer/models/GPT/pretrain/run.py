"""GPT pre-training with an explicit loop (no Engine / Module objects) on the ``distributed.apis`` building blocks
(reference examples/transformer/models/GPT/pretrain/run.py).

    python examples/transformer/models/GPT/pretrain/run.py -c examples/transformer/models/GPT/pretrain/configs/pretrain_gpt_345M_single_card.yaml
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/transformer/models/GPT/pretrain/run.py -c .../pretrain_gpt_6.7B_sharding16.yaml
"""
import os
import sys
import time

__dir__ = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(__dir__, "..", "..", "..", "..", "..")))
sys.path.insert(0, os.path.abspath(os.path.join(__dir__, "..", "..", "..")))

import torch  # noqa: E402

import impls  # noqa: E402
from paddlefleetx_b200.distributed.apis import amp as amp_api, env, io, strategy  # noqa: E402
from paddlefleetx_b200.utils.log import logger  # noqa: E402
from utils import components as cpn  # noqa: E402
from utils import config as cfg_utils  # noqa: E402


def main():
    args = cfg_utils.parse_args()
    config = cfg_utils.get_config(args.config, overrides=args.override, show=False)
    if env.world_size() > 1:
        env.init_dist_env(config)
    env.set_seed(config.Global.seed)
    device = torch.device("cuda", torch.cuda.current_device()) if (config.Global.device != "cpu" and torch.cuda.is_available()) else torch.device("cpu")

    model, loss_fn = impls.build_model(config)
    lr = cpn.build_lr_scheduler(config.Optimizer.lr)
    optimizer = cpn.build_optimizer(config.Optimizer, model, lr, multi_precision=config.Engine.mix_precision.get("enable", False),
                                    dist_config=config.Distributed, amp_config=config.Engine.mix_precision)
    mp = config.Engine.mix_precision
    scaler = amp_api.GradScaler(init_loss_scaling=mp.get("scale_loss", 32768.0)) if (mp.get("enable") and mp.get("dtype") == "float16") else None
    if env.world_size() > 1:
        model, optimizer, scaler = strategy.wrap_with_fleet(config.Distributed, model, optimizer, scaler)
    if config.Engine.save_load.get("ckpt_dir"):
        rec = io.load(config.Engine.save_load.ckpt_dir, model, optimizer, "train", scaler=scaler)
        start_step = rec.get("step", 0)
    else:
        rec, start_step = None, 0

    train_loader = cpn.build_dataloader(config.Data, "Train")
    valid_loader = cpn.build_dataloader(config.Data, "Eval") if config.Data.get("Eval") else None
    profiler = cpn.build_profiler(config.get("Profiler"))
    eng = config.Engine
    tokens_per_step = config.Global.global_batch_size * config.Data.Train.dataset.max_seq_len
    t0, losses = time.time(), []
    for step, batch in enumerate(train_loader):
        if step < start_step:
            continue
        if step >= eng.max_steps:
            break
        if rec is not None:           # first step after a resume: dropout / routing noise continues from the checkpointed streams
            io.restore_rng(rec)
            rec = None
        batch = [t.to(device, non_blocking=True) for t in batch]
        losses.append(impls.fit_impl(config, batch, model, loss_fn, optimizer, scaler))
        lr.step()
        if (step + 1) % eng.logging_freq == 0:
            if device.type == "cuda":
                torch.cuda.synchronize()
            dt = (time.time() - t0) / eng.logging_freq
            logger.train("[train] step: %d/%d, loss: %.9f, avg_batch_cost: %.5f sec, ips_total: %.0f tokens/s, learning rate: %.5e" % (
                step + 1, eng.max_steps, float(sum(float(l) for l in losses) / len(losses)), dt, tokens_per_step / dt, optimizer.get_lr()))
            t0, losses = time.time(), []
        if valid_loader is not None and eng.eval_freq > 0 and (step + 1) % eng.eval_freq == 0:
            ev = [float(impls.eval_impl(config, [t.to(device) for t in b], model, loss_fn)) for _, b in zip(range(eng.eval_iters), valid_loader)]
            logger.eval("[eval] step: %d, loss: %.9f" % (step + 1, sum(ev) / max(len(ev), 1)))
        if eng.save_load.save_steps > 0 and (step + 1) % eng.save_load.save_steps == 0:
            io.save(eng.save_load.output_dir, model, optimizer, step=step + 1, epoch=0, scaler=scaler)
        if profiler is not None:
            profiler.step()
    cpn.profiler_done(profiler)


if __name__ == "__main__":
    main()

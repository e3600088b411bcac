"""python tools/auto.py -c <yaml> [-o k=v]  —  "auto-parallel" entry (reference tools/auto.py:40-69): same YAML surface, executed by
the eager hybrid engine on the mesh derived from the degrees; ``-o Tuning.enable=True`` prints the planner's ranked layouts."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

from paddlefleetx_b200.core.engine.auto_engine import AutoEngine  # noqa: E402
from paddlefleetx_b200.data import build_dataloader  # noqa: E402
from paddlefleetx_b200.distributed.apis import env  # noqa: E402
from paddlefleetx_b200.models import build_module  # noqa: E402
from paddlefleetx_b200.utils import config  # noqa: E402


def main(argv=None):
    args = config.parse_args(argv)
    env.init_process_group("gpu")
    cfg = config.get_auto_config(args.config, overrides=args.override, show=False)
    env.init_dist_env(cfg)
    env.set_seed(cfg.Global.seed)
    module = build_module(cfg)
    config.print_config(cfg)
    if cfg.Distributed.get("plan") is not None:
        print(f"[auto_layout] {cfg.Distributed.plan.describe}")
    engine = AutoEngine(configs=cfg, module=module)
    train_loader = build_dataloader(cfg.Data, "Train")
    if cfg.get("Tuning", {}).get("enable", False):
        for row in engine.tune(train_loader)[:8]:
            print(row.get("describe", row) if isinstance(row, dict) else row)
        return engine
    eval_loader = build_dataloader(cfg.Data, "Eval") if cfg.Engine.eval_freq and cfg.Engine.eval_freq > 0 and "Eval" in cfg.Data else None
    if cfg.Engine.save_load.ckpt_dir is not None:
        engine.load()
    engine.fit(train_data_loader=train_loader, valid_data_loader=eval_loader, epoch=cfg.Engine.num_train_epochs)
    return engine


if __name__ == "__main__":
    main()

"""python tools/eval.py -c <yaml> [-o k=v]  —  offline evaluation (reference tools/eval.py:34-54)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

from paddlefleetx_b200.core import EagerEngine  # noqa: E402
from paddlefleetx_b200.data import build_dataloader  # noqa: E402
from paddlefleetx_b200.distributed.apis import env  # noqa: E402
from paddlefleetx_b200.models import build_module  # noqa: E402
from paddlefleetx_b200.utils import config  # noqa: E402


def main(argv=None):
    args = config.parse_args(argv)
    env.init_process_group("gpu")
    cfg = config.get_config(args.config, overrides=args.override, show=False)
    env.init_dist_env(cfg)
    env.set_seed(cfg.Global.seed)
    module = build_module(cfg)
    config.print_config(cfg)
    loader = build_dataloader(cfg.Data, "Eval")
    engine = EagerEngine(configs=cfg, module=module, mode="eval")
    if cfg.Engine.save_load.ckpt_dir is not None:
        engine.load()
    # one pass: the checkpoint is fixed, so the reference's ``epoch=num_train_epochs`` (tools/eval.py:53) just repeats the same numbers
    engine.evaluate(valid_data_loader=loader, epoch=int(cfg.Engine.get("eval_epochs", 1)))
    return engine


if __name__ == "__main__":
    main()

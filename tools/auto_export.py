"""python tools/auto_export.py -c <yaml>  —  export from an auto-parallel config (reference tools/auto_export.py:33-53)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

from paddlefleetx_b200.core.engine.auto_engine import AutoEngine  # noqa: E402
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
    engine = AutoEngine(configs=cfg, module=module, mode="export")
    if cfg.Engine.save_load.ckpt_dir is not None:
        engine.load()
    engine.export_from_prog()
    return engine


if __name__ == "__main__":
    main()

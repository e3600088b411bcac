"""python tools/inference.py -c <yaml> [-o k=v]  —  run the exported model over the Test dataloader (reference tools/inference.py:37-59)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

import numpy as np  # noqa: E402

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
    engine = EagerEngine(configs=cfg, module=module, mode="inference")
    loader = build_dataloader(cfg.Data, "Test")
    outs = []
    for it, batch in enumerate(loader):
        arrays = [np.asarray(t) for t in (batch if isinstance(batch, (list, tuple)) else [batch])]
        n_in = len(engine._module.input_spec())
        outs.append(module.inference_end(engine.inference(arrays[:n_in])))
        if it + 1 >= cfg.Engine.test_iters:
            break
    return outs


if __name__ == "__main__":
    main()

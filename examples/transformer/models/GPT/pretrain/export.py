"""Export the pre-trained GPT for serving (reference examples/transformer/models/GPT/pretrain/export.py)."""
import os
import sys

__dir__ = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(__dir__, "..", "..", "..", "..", "..")))
sys.path.insert(0, os.path.abspath(os.path.join(__dir__, "..", "..", "..")))

import impls  # noqa: E402
from paddlefleetx_b200.distributed.apis import env, io  # noqa: E402
from paddlefleetx_b200.utils.export import export_inference_model  # noqa: E402
from utils import config as cfg_utils  # noqa: E402


def main():
    args = cfg_utils.parse_args()
    config = cfg_utils.get_config(args.config, overrides=args.override)
    if env.world_size() > 1:
        env.init_dist_env(config)
    model, _ = impls.build_model(config)
    if config.Engine.save_load.get("ckpt_dir"):
        io.load(config.Engine.save_load.ckpt_dir, model, None, "eval")
    model.eval()
    seq = config.Data.Train.dataset.max_seq_len
    spec = [{"name": "tokens", "shape": [None, seq], "dtype": "int64"}, {"name": "ids", "shape": [None, seq], "dtype": "int64"}]
    out = os.path.join(config.Engine.save_load.output_dir, "rank_0")
    export_inference_model(model, spec, out, "model", configs=config)
    print(f"exported to {out}")


if __name__ == "__main__":
    main()

"""Generate text from a prompt (reference examples/transformer/models/GPT/generation/run.py)."""
import os
import sys

__dir__ = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(__dir__, "..", "..", "..", "..", "..")))
sys.path.insert(0, os.path.abspath(os.path.join(__dir__, "..", "..", "..")))

import torch  # noqa: E402

import impls  # noqa: E402
from paddlefleetx_b200.distributed.apis import env, io  # noqa: E402
from utils import config as cfg_utils  # noqa: E402


def main():
    args = cfg_utils.parse_args()
    config = cfg_utils.get_config(args.config, overrides=args.override)
    if env.world_size() > 1:
        env.init_dist_env(config)
    env.set_seed(config.Global.seed)
    module = impls.build_module(config)
    ckpt = config.Engine.save_load.get("ckpt_dir")
    if ckpt:
        io.load(ckpt, module.model, None, "eval")
    module.model.eval()
    prompt = os.environ.get("PROMPT", "Hi, GPT2. Tell me who Jack Ma is.")
    with torch.no_grad():
        out = impls.generate(module, prompt)
    print("Prompt:", prompt)
    print("Generation:", out[0])


if __name__ == "__main__":
    main()

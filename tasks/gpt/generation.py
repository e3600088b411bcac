"""python tasks/gpt/generation.py -c <generation yaml> [-o Generation.text="..."]  —  eager text generation
(reference tasks/gpt/generation.py:34-63): build ``GPTGenerationModule``, load ``model.pdparams`` (if ckpt_dir is set), generate."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

import torch  # noqa: E402

from paddlefleetx_b200.distributed.apis import env  # noqa: E402
from paddlefleetx_b200.models import build_module  # noqa: E402
from paddlefleetx_b200.utils import config  # noqa: E402
from paddlefleetx_b200.utils.log import logger  # noqa: E402


def main(argv=None):
    args = config.parse_args(argv)
    env.init_process_group("gpu")
    cfg = config.get_config(args.config, overrides=args.override, show=False)
    env.init_dist_env(cfg)
    env.set_seed(cfg.Global.seed)
    module = build_module(cfg)
    ckpt_dir = cfg.Engine.save_load.ckpt_dir
    if ckpt_dir is not None:
        path = os.path.join(ckpt_dir, "model.pdparams")
        state = torch.load(path, map_location="cpu", weights_only=False)
        own = module.model.state_dict()
        module.model.load_state_dict({k: v.to(own[k].dtype) for k, v in state.items() if k in own}, strict=False)
        logger.info(f"load model from {path}")
    text = cfg.Generation.get("text", "Hi, GPT2. Tell me who Jack Ma is.")
    for out in module.generate(text):
        print("Prompt:", text)
        print("Generation:", out[len(text):] if isinstance(text, str) else out)
    return module


if __name__ == "__main__":
    main()

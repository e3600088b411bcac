"""Serve an exported generation model (reference .../generation/inference.py)."""
import os
import sys

__dir__ = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(__dir__, "..", "..", "..", "..", "..")))
sys.path.insert(0, os.path.abspath(os.path.join(__dir__, "..", "..", "..")))

import numpy as np  # noqa: E402

from paddlefleetx_b200.core.engine.inference_engine import InferenceEngine  # noqa: E402
from paddlefleetx_b200.data.tokenizers.gpt_tokenizer import GPTTokenizer  # noqa: E402
from utils import config as cfg_utils  # noqa: E402


def main():
    args = cfg_utils.parse_args()
    config = cfg_utils.get_config(args.config, overrides=args.override)
    eng = InferenceEngine(config.Inference.model_dir, config.Inference.get("mp_degree", 1))
    tok = GPTTokenizer.from_pretrained("gpt2")
    prompt = os.environ.get("PROMPT", "Hi, GPT2. Tell me who Jack Ma is.")
    ids = np.array([tok.encode(prompt)], dtype=np.int64)
    out = eng.predict([ids])
    seq = list(out.values())[0] if isinstance(out, dict) else out[0]
    print("Prompt:", prompt)
    print("Generation:", tok.decode(np.asarray(seq)[0].tolist()))


if __name__ == "__main__":
    main()

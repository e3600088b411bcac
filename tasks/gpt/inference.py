"""python tasks/gpt/inference.py --model_dir <export dir> [--mp_degree N] [--text "..."]  —  stand-alone ``InferenceEngine``
driver: tokenise -> predict -> decode (reference tasks/gpt/inference.py, projects/gpt/inference.py:42-66)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

import numpy as np  # noqa: E402

from paddlefleetx_b200.core.engine.inference_engine import InferenceEngine  # noqa: E402
from paddlefleetx_b200.data.tokenizers import GPTTokenizer  # noqa: E402


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--model_dir", default="./output")
    p.add_argument("--mp_degree", type=int, default=1)
    p.add_argument("--vocab_dir", default="gpt2")
    p.add_argument("--text", default="Hi, GPT2. Tell me who Jack Ma is.")
    a = p.parse_args(argv)
    try:
        tok = GPTTokenizer.from_pretrained(a.vocab_dir)
    except FileNotFoundError:
        tok = GPTTokenizer.byte_fallback()
    engine = InferenceEngine(a.model_dir, a.mp_degree)
    ids = np.asarray([tok.encode(a.text)], dtype=np.int64)
    outs = engine.predict([ids])
    first = next(iter(outs.values()))
    print("Prompt:", a.text)
    if first.ndim == 2 and np.issubdtype(first.dtype, np.integer):          # a generation export: token ids
        print("Generation:", tok.decode([int(t) for t in first[0]], skip_special_tokens=True))
    else:                                                                    # a pre-training / QAT export: logits [batch, seq, vocab]
        nxt = int(np.asarray(first)[0, -1].argmax())
        print(f"Logits {tuple(first.shape)}; most likely next token: {nxt} {tok.decode([nxt])!r} (export a generation recipe to sample text)")
    return outs


if __name__ == "__main__":
    main()

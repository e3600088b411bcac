"""Serve an exported ERNIE model (reference projects/ernie/inference.py): feeds ``input_ids / token_type_ids`` of a text (or of
random ids when no vocabulary is available) to the ``InferenceEngine`` and prints the output shapes."""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

import numpy as np  # noqa: E402

from paddlefleetx_b200.core.engine.inference_engine import InferenceEngine  # noqa: E402


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--model_dir", default="./output")
    p.add_argument("--mp_degree", type=int, default=1)
    p.add_argument("--seq_len", type=int, default=128)
    p.add_argument("--batch", type=int, default=1)
    p.add_argument("--vocab_dir", default=None, help="directory with the ERNIE vocab.txt; without it random ids are fed")
    p.add_argument("--text", default="Hello, my dog is cute")
    a = p.parse_args(argv)
    engine = InferenceEngine(a.model_dir, a.mp_degree)
    vocab = engine.recipe.get("Model", {}).get("vocab_size", 40000)
    if a.vocab_dir:
        from paddlefleetx_b200.data.tokenizers import get_ernie_tokenizer

        enc = get_ernie_tokenizer(a.vocab_dir)([a.text] * a.batch, padding="max_length", truncation=True, max_length=a.seq_len, return_tensors="np")
        ids, seg = enc["input_ids"].astype(np.int64), enc["token_type_ids"].astype(np.int64)
    else:
        rng = np.random.RandomState(0)
        ids = rng.randint(1, vocab - 1, size=(a.batch, a.seq_len)).astype(np.int64)
        seg = np.zeros_like(ids)
    outs = engine.predict([ids, seg])
    for k, v in outs.items():
        print(k, tuple(v.shape), v.dtype)
    return outs


if __name__ == "__main__":
    main()

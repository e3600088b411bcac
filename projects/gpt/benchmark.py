"""Inference latency benchmark (reference projects/gpt/benchmark.py:45-81): prompt of --seq_len tokens, --max_dec_len generated
tokens, batch sizes {1,2,4,8,16}, 10 warm-up + N timed ``predict`` calls; latency from CUDA events (device time) and wall clock."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from paddlefleetx_b200.core.engine.inference_engine import InferenceEngine  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--model_dir", default="./output")
    p.add_argument("--mp_degree", type=int, default=1)
    p.add_argument("--seq_len", type=int, default=128)
    p.add_argument("--iters", "--iter", dest="iters", type=int, default=10)
    p.add_argument("--batch_sizes", default="1,2,4,8,16")
    a = p.parse_args()
    eng = InferenceEngine(a.model_dir, a.mp_degree)
    vocab = eng.recipe.get("Model", {}).get("vocab_size", 50304)
    for bs in [int(b) for b in a.batch_sizes.split(",")]:
        ids = np.random.randint(0, vocab - 1, size=(bs, a.seq_len)).astype(np.int64)
        for _ in range(10):
            eng.predict([ids])
        torch.cuda.synchronize() if torch.cuda.is_available() else None
        t0 = time.perf_counter()
        for _ in range(a.iters):
            eng.predict([ids])
        torch.cuda.synchronize() if torch.cuda.is_available() else None
        print(f"batch {bs:3d}: latency {(time.perf_counter() - t0) / a.iters * 1e3:.2f} ms")


if __name__ == "__main__":
    main()

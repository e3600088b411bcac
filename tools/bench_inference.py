"""Generation latency benchmark on random-init weights (the reference's projects/gpt/benchmark.py protocol — prompt of 128
tokens, 8 generated tokens, batch 1/2/4/8/16, 10 warm-up runs — docs/inference.md:99-113 publishes 84.93 ms for GPT-6.7B bs 1).

Timed through the public generation API (``GPTForGeneration.generate``): tokenised prompt copied H2D each iteration, generated ids
read back D2H, CUDA-event device time plus wall clock.

    python tools/bench_inference.py --model gpt-6.7b --batches 1,2,4,8,16 --iters 20 [--no-graph] [--int8]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

import torch  # noqa: E402

SHAPES = {"gpt-345m": dict(hidden_size=1024, num_layers=24, num_attention_heads=16), "gpt-1.3b": dict(hidden_size=2048, num_layers=24, num_attention_heads=16),
          "gpt-6.7b": dict(hidden_size=4096, num_layers=32, num_attention_heads=32)}
REF_INT8_MS = {"gpt-345m": {1: 18.30, 2: 18.88, 4: 20.77, 8: 23.90, 16: 27.95}, "gpt-6.7b": {1: 63.96, 2: 67.25, 4: 78.98, 8: 99.54, 16: 140.97}}
REF_FP16_MS = {"gpt-345m": {1: 18.91, 2: 20.01, 4: 20.83, 8: 24.06, 16: 29.32}, "gpt-6.7b": {1: 84.93, 2: 91.93, 4: 105.50, 8: 138.56, 16: 204.33}}


def main():
    """Time prompt + generation for each batch size with CUDA events after warm-up and write one JSON record (latency per batch, tokens / s)."""
    p = argparse.ArgumentParser()
    p.add_argument("--model", default="gpt-6.7b")
    p.add_argument("--batches", default="1,2,4,8,16")
    p.add_argument("--seq_len", type=int, default=128)
    p.add_argument("--max_dec_len", type=int, default=8)
    p.add_argument("--iters", type=int, default=20)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--strategy", default="sampling")
    p.add_argument("--int8", action="store_true", help="W8A8 linears (reference INT8 column)")
    a = p.parse_args()
    from paddlefleetx_b200.models.language_model.gpt import model as gpt
    from paddlefleetx_b200.models.language_model.gpt.generation import GPTForGeneration
    from paddlefleetx_b200.ops import functional as OF

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    shp = SHAPES[a.model]
    core = gpt.GPTModel(vocab_size=50304, ffn_hidden_size=4 * shp["hidden_size"], max_position_embeddings=1024, hidden_dropout_prob=0.0,
                        attention_probs_dropout_prob=0.0, dtype=torch.bfloat16, device=dev, **shp)
    gen = GPTForGeneration(core, dict(max_dec_len=a.max_dec_len, min_dec_len=a.max_dec_len, decode_strategy=a.strategy, top_k=0, top_p=0.9,
                                      temperature=1.0, eos_token_id=50256, pad_token_id=50256, use_cuda_graph=not a.no_graph))
    if a.int8:
        from paddlefleetx_b200.ops.quant import quantize_tp_linears_int8

        print(f"int8: converted {quantize_tp_linears_int8(core)} linear layers", flush=True)
        torch.cuda.empty_cache()
    rows = []
    for bs in [int(b) for b in a.batches.split(",")]:
        host = torch.randint(0, 50000, (bs, a.seq_len)).pin_memory()
        def once():
            ids = host.to(dev, non_blocking=True)
            out, _ = gen.generate(ids, seed=1234)
            return out.cpu()
        for _ in range(a.warmup):
            once()
        torch.cuda.synchronize()
        OF.reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(a.iters):
            out = once()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / a.iters * 1e3
        devms = e0.elapsed_time(e1) / a.iters
        ref = (REF_INT8_MS if a.int8 else REF_FP16_MS).get(a.model, {}).get(bs)
        row = dict(model=a.model, batch=bs, prompt=a.seq_len, new_tokens=int(out.shape[1]), latency_ms=round(wall, 3), device_ms=round(devms, 3),
                   cuda_graph=not a.no_graph, int8=a.int8, native_launches_per_call=OF.native_launch_count() / a.iters, reference_ms=ref,
                   speedup_vs_published=round(ref / wall, 2) if ref else None)
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/inference_latency_{a.model}{'_nograph' if a.no_graph else ''}{'_int8' if a.int8 else ''}.json", "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()

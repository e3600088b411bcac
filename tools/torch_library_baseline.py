"""``bench.py --impl torch_library``: the headline model/config on STOCK PyTorch — cuBLAS linears, library SDPA (cuDNN / flash),
``torch.optim.AdamW(fused=True)`` on fp32 parameters under bf16 autocast, NCCL DDP (bf16-compressed gradient all-reduce) for N > 1.

This arm exists because the reference (PaddleFleetX on Paddle) cannot be installed offline; it is NOT the reference and is labelled
``"impl": "torch_library"``.  None of this repository's kernels, models or engine are on its path: it answers "what does the same
training step cost when every hot op is a library call on the same box", which is the number our hand-written path has to beat.
"""
from __future__ import annotations

import json
import os

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F


class _Block(nn.Module):
    def __init__(self, h: int, heads: int, p: float):
        super().__init__()
        self.ln1, self.ln2 = nn.LayerNorm(h), nn.LayerNorm(h)
        self.qkv, self.proj = nn.Linear(h, 3 * h), nn.Linear(h, h)
        self.fc1, self.fc2 = nn.Linear(h, 4 * h), nn.Linear(4 * h, h)
        self.heads, self.p = heads, p

    def forward(self, x):
        b, s, h = x.shape
        q, k, v = self.qkv(self.ln1(x)).view(b, s, 3, self.heads, h // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, is_causal=True, dropout_p=self.p if self.training else 0.0)
        x = x + F.dropout(self.proj(a.transpose(1, 2).reshape(b, s, h)), self.p, self.training)
        return x + F.dropout(self.fc2(F.gelu(self.fc1(self.ln2(x)), approximate="tanh")), self.p, self.training)


class _GPT(nn.Module):
    def __init__(self, vocab: int, h: int, layers: int, heads: int, seq: int, p: float = 0.1):
        super().__init__()
        self.wte, self.wpe = nn.Embedding(vocab, h), nn.Embedding(seq, h)
        self.blocks = nn.ModuleList(_Block(h, heads, p) for _ in range(layers))
        self.ln_f = nn.LayerNorm(h)
        self.p = p
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                nn.init.normal_(m.weight, 0.0, 0.02)

    def forward(self, tokens, pos, labels, mask):
        x = F.dropout(self.wte(tokens) + self.wpe(pos), self.p, self.training)
        for blk in self.blocks:
            x = blk(x)
        logits = F.linear(self.ln_f(x), self.wte.weight)
        loss = F.cross_entropy(logits.float().view(-1, logits.shape[-1]), labels.reshape(-1), reduction="none")
        return (loss * mask.reshape(-1)).sum() / mask.sum()


def run(args, spec, ClockSampler) -> int:
    """Train the headline config on stock PyTorch under the same timing rules as our arm and print the same JSON line with ``impl: torch_library``."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dev = torch.device("cuda", torch.cuda.current_device())
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    vocab, seq, local = 50304, args.seq_len, args.local_batch
    try:
        torch.manual_seed(1234)
        with torch.device(dev):
            model = _GPT(vocab, spec["hidden"], args.layers or spec["layers"], spec["heads"], seq)
        net = model
        if world > 1:
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks

            net = nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], bucket_cap_mb=512, gradient_as_bucket_view=True, static_graph=True)
            net.register_comm_hook(None, default_hooks.bf16_compress_hook)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.999), weight_decay=0.01, fused=True)
        gen = torch.Generator().manual_seed(1234 + rank)

        def host_batch():
            toks = torch.randint(0, vocab, (local, seq + 1), generator=gen, dtype=torch.int64)
            pos = torch.arange(seq, dtype=torch.int64).unsqueeze(0).expand(local, seq).contiguous()
            return [t.pin_memory() for t in (toks[:, :-1].contiguous(), pos, toks[:, 1:].contiguous(), torch.ones(local, seq))]

        pool = [host_batch() for _ in range(4)]
        dev_pool = [[t.to(dev) for t in b] for b in pool]

        def step(batch):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = net(*batch)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0, foreach=True)
            opt.step()
            opt.zero_grad(set_to_none=False)
            return loss

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        for i in range(args.warmup):
            step(dev_pool[i % 4])
        barrier()
        sampler = ClockSampler(dev.index) if rank == 0 else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(args.steps):
            loss = step(dev_pool[i % 4])
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(args.steps):
            _ = step([x.to(dev, non_blocking=True) for x in pool[i % 4]]).item()
        f1.record()
        barrier()
        t2 = torch.tensor([f0.elapsed_time(f1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        clocks = sampler.stop() if sampler else None
        if rank == 0:
            tokens = local * world * seq * args.steps
            print(json.dumps({
                "impl": "torch_library",
                "metric": f"GPT-3 {args.model.split('-')[1].upper()} pre-training tokens/sec (whole job, device-timed, max over ranks)",
                "value": tokens / (float(t) / 1e3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": float(t) / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic (uniform random tokens, random-init weights)",
                "config": {"model": args.model if not args.layers else f"{args.model}-DEBUG-{args.layers}layers(INVALID)", "global_batch": local * world,
                           "seq_len": seq, "parallelism": "single" if world == 1 else f"ddp{world} (bf16-compressed all-reduce)",
                           "stack": "torch.nn (cuBLAS), F.scaled_dot_product_attention, torch.optim.AdamW(fused) on fp32 params under bf16 autocast"},
                "gpu_launches": 0, "clocks": clocks,
                "e2e": {"value": tokens / (float(t2) / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": sum(x.numel() * x.element_size() for x in pool[0]),
                        "d2h_bytes_per_step": 4},
                "final_loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    except torch.OutOfMemoryError as e:
        if rank == 0:
            print(json.dumps({"impl": "torch_library", "unavailable": f"out of memory in the stock-PyTorch arm: {str(e)[:160]}"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0

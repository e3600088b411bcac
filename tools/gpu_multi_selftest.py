"""Multi-GPU checks of the peer-memory kernels (run with torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_selftest.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def relerr(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def timed(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from paddlefleetx_b200.ops import _native
    from paddlefleetx_b200.parallel import fused_tp
    from paddlefleetx_b200.parallel.symmetric_memory import SymmetricAllocator
    from paddlefleetx_b200.parallel.topology import HybridCommunicateGroup

    lib = _native.require()
    hcg = HybridCommunicateGroup(mp=world)
    grp = hcg.get_model_parallel_group()
    res = {}

    def report(name, **kw):
        res[name] = kw
        if rank == 0:
            print("RESULT " + json.dumps(dict(check=name, **kw)), flush=True)

    # ---------------------------------------------------------------- barrier + raw P2P
    sm = SymmetricAllocator(grp)
    n = 1 << 24
    buf = sm.alloc_tensor(n * world, torch.bfloat16)
    torch.manual_seed(100 + rank)
    local = (torch.randn(n * world, device="cuda") * 0.1).bfloat16()
    buf.copy_(local)
    torch.cuda.synchronize(); sm.barrier()
    ptrs = sm.peer_ptrs(buf)
    # reference reduce-scatter via NCCL
    ref = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    dist.reduce_scatter_tensor(ref, local.clone())
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    lib.p2p_reduce_scatter(ptrs, out, rank, 1, False, 1.0, 128)
    torch.cuda.synchronize(); sm.barrier()
    report("p2p_reduce_scatter", err=relerr(out, ref), ok=relerr(out, ref) < 1e-2)

    def rs_p2p():
        sm.barrier(); lib.p2p_reduce_scatter(ptrs, out, rank, 1, False, 1.0, 128); sm.barrier()
    t_p2p = timed(rs_p2p)
    scratch = local.clone()
    t_nccl = timed(lambda: dist.reduce_scatter_tensor(ref, scratch))
    moved = n * (world - 1) * 2
    report("p2p_reduce_scatter_perf", ms=t_p2p, nccl_ms=t_nccl, gbs_in=moved / t_p2p / 1e6, nccl_gbs_in=moved / t_nccl / 1e6, ok=True)

    # push all-gather
    shard = (torch.arange(n, device="cuda") % 251 + rank).bfloat16()
    sm.barrier()
    lib.p2p_all_gather(ptrs, shard, rank, 32)
    torch.cuda.synchronize(); sm.barrier()
    ok = True
    for r in range(world):
        want = (torch.arange(n, device="cuda") % 251 + r).bfloat16()
        ok &= bool(torch.equal(buf[r * n:(r + 1) * n], want))
    report("p2p_all_gather", ok=ok)
    t_ag = timed(lambda: (sm.barrier(), lib.p2p_all_gather(ptrs, shard, rank, 32), sm.barrier()))
    gath = torch.empty(n * world, dtype=torch.bfloat16, device="cuda")
    t_ag_nccl = timed(lambda: dist.all_gather_into_tensor(gath, shard))
    report("p2p_all_gather_perf", ms=t_ag, nccl_ms=t_ag_nccl, gbs_out=moved / t_ag / 1e6, nccl_gbs=moved / t_ag_nccl / 1e6, ok=True)

    # ---------------------------------------------------------------- fused GEMM -> reduce-scatter
    torch.manual_seed(7)
    M, N, K = 8192, 4096, 4096 // world * world
    kl = K // world
    A = (torch.randn(M, K, device="cuda") * 0.05).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    a_loc, w_loc = A[:, rank * kl:(rank + 1) * kl].contiguous(), W[:, rank * kl:(rank + 1) * kl].contiguous()
    full = A.float() @ W.float().t()
    want = full[rank * (M // world):(rank + 1) * (M // world)]
    got = fused_tp.gemm_rs(a_loc, w_loc, grp)
    torch.cuda.synchronize()
    report("gemm_rs", err=relerr(got, want), ok=relerr(got, want) < 2e-2)
    t_f = timed(lambda: fused_tp.gemm_rs(a_loc, w_loc, grp))
    part = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    o2 = torch.empty(M // world, N, dtype=torch.bfloat16, device="cuda")

    def baseline_rs():
        torch.matmul(a_loc, w_loc.t(), out=part)
        dist.reduce_scatter_tensor(o2, part)
    t_b = timed(baseline_rs)
    t_g = timed(lambda: torch.matmul(a_loc, w_loc.t(), out=part))
    report("gemm_rs_perf", fused_ms=t_f, cublas_plus_nccl_ms=t_b, gemm_only_ms=t_g, ok=True)

    # ---------------------------------------------------------------- fused all-gather -> GEMM
    rows = M // world
    X = (torch.randn(M, K, device="cuda") * 0.05).bfloat16()        # same on all ranks (same seed)
    Wc = (torch.randn(N // world * world, K, device="cuda") * 0.05).bfloat16()
    nl = Wc.shape[0] // world
    w_col = Wc[rank * nl:(rank + 1) * nl].contiguous()
    x_shard = X[rank * rows:(rank + 1) * rows].contiguous()
    y, gathered = fused_tp.ag_gemm(x_shard, w_col, None, grp)
    torch.cuda.synchronize()
    want = X.float() @ w_col.float().t()
    report("ag_gemm", err=relerr(y, want), gathered_ok=bool(torch.equal(gathered, X)), ok=relerr(y, want) < 2e-2 and bool(torch.equal(gathered, X)))
    t_f = timed(lambda: fused_tp.ag_gemm(x_shard, w_col, None, grp))
    xg = torch.empty(M, K, dtype=torch.bfloat16, device="cuda")
    yo = torch.empty(M, nl, dtype=torch.bfloat16, device="cuda")

    def baseline_ag():
        dist.all_gather_into_tensor(xg, x_shard)
        torch.matmul(xg, w_col.t(), out=yo)
    t_b = timed(baseline_ag)
    t_g = timed(lambda: torch.matmul(xg, w_col.t(), out=yo))
    report("ag_gemm_perf", fused_ms=t_f, nccl_plus_cublas_ms=t_b, gemm_only_ms=t_g, ok=True)

    # ---------------------------------------------------------------- autograd parity of the fused SP linears
    from paddlefleetx_b200.ops import functional as OF
    from paddlefleetx_b200.parallel import comm_ops as C

    s, b, h = 1024, 2, 1024
    torch.manual_seed(3)
    xs = (torch.randn(s // world, b, h, device="cuda") * 0.5).bfloat16().requires_grad_(True)
    wq = (torch.randn(3 * h // world, h, device="cuda") * 0.05).bfloat16().requires_grad_(True)
    bq = (torch.randn(3 * h // world, device="cuda") * 0.05).bfloat16().requires_grad_(True)
    wo = (torch.randn(h, 3 * h // world, device="cuda") * 0.05).bfloat16().requires_grad_(True)
    go = (torch.randn(s // world, b, h, device="cuda") * 0.1).bfloat16()

    def run(fused):
        for t in (xs, wq, bq, wo):
            t.grad = None
        if fused:
            mid = fused_tp.all_gather_linear(xs, wq, bq, grp)
            out = fused_tp.linear_reduce_scatter(mid, wo, grp)
        else:
            mid = OF.linear(C.all_gather_seq(xs, grp), wq, bq)
            out = C.reduce_scatter_seq(OF.linear(mid, wo, None), grp)
        out.backward(go)
        return [out.detach().clone()] + [t.grad.detach().clone() for t in (xs, wq, bq, wo)]

    ref_out = run(False)
    fus_out = run(True)
    errs = [relerr(a, b_) for a, b_ in zip(fus_out, ref_out)]
    report("fused_sp_linear_autograd", errs=errs, ok=max(errs) < 3e-2)

    # ---------------------------------------------------------------- flat optimizer: P2P path == NCCL path
    from paddlefleetx_b200.optims import ClipGradByGlobalNorm, FusedAdamW

    hc2 = HybridCommunicateGroup(sharding=world)
    torch.manual_seed(11)
    def make():
        m = torch.nn.Sequential(torch.nn.Linear(1024, 2048), torch.nn.Linear(2048, 1024)).cuda().bfloat16()
        return m
    m1, m2 = make(), make()
    m2.load_state_dict(m1.state_dict())
    o1 = FusedAdamW(1e-2, named_parameters=list(m1.named_parameters()), multi_precision=True, hcg=hc2, grad_clip=ClipGradByGlobalNorm(1.0, hc2))
    o2_ = FusedAdamW(1e-2, named_parameters=list(m2.named_parameters()), multi_precision=True, hcg=hc2, use_p2p=True,
                     grad_clip=ClipGradByGlobalNorm(1.0, hc2))
    for step in range(3):
        # the SAME batch on every rank: the cross-rank gradient sum is then exact in bf16 (x world), so the NCCL ring (bf16
        # accumulation) and the peer-memory pull (fp32 accumulation) must agree bit for bit at any world size
        torch.manual_seed(50 + 10 * step)
        xin = torch.randn(64, 1024, device="cuda").bfloat16()
        for m, o in ((m1, o1), (m2, o2_)):
            m(xin).float().pow(2).mean().backward()
            o.step(); o.clear_grad()
    errs = [relerr(a, b_) for a, b_ in zip(m2.parameters(), m1.parameters())]
    report("zero_p2p_vs_nccl", errs=errs, ok=max(errs) < 5e-3)

    # ---------------------------------------------------------------- MoE: peer-memory dispatch/combine == NCCL all-to-all path
    from paddlefleetx_b200.models.language_model.moe.moe_layer import ExpertLayer, MoELayer

    hc3 = HybridCommunicateGroup(dp=world)
    mg = hc3.get_moe_group()
    hm, e_local, tok = 1024, 2, 8192
    torch.manual_seed(7 + rank)
    def make_moe(fused):
        torch.manual_seed(7 + rank)
        ex = [ExpertLayer(hm, 4 * hm, dtype=torch.bfloat16, device="cuda") for _ in range(e_local)]
        return MoELayer(hm, ex, gate={"type": "naive", "top_k": 2}, moe_group=mg, dtype=torch.bfloat16, device="cuda", fused_p2p=fused)
    l_ref, l_p2p = make_moe(False), make_moe(True)
    l_p2p.load_state_dict(l_ref.state_dict())
    torch.manual_seed(99 + rank)
    xin = (torch.randn(tok, hm, device="cuda") * 0.5).bfloat16()
    gout = (torch.randn(tok, hm, device="cuda") * 0.1).bfloat16()
    outs = []
    for layer in (l_ref, l_p2p):
        xi = xin.clone().requires_grad_(True)
        y = layer(xi)
        y.backward(gout)
        outs.append([y.detach(), xi.grad.detach()] + [p.grad.detach().float() for p in layer.parameters()])
    errs = [relerr(a, b_) for a, b_ in zip(outs[1], outs[0])]
    def moe_step(layer):
        def f():
            xi = xin.clone().requires_grad_(True)
            layer(xi).backward(gout)
        return f
    t_ref, t_p2p = timed(moe_step(l_ref), iters=5), timed(moe_step(l_p2p), iters=5)
    report("moe_p2p_dispatch_combine", errs=[round(e, 5) for e in errs], ok=max(errs) < 3e-2, ms_nccl=t_ref, ms_p2p=t_p2p,
           shape=dict(tokens=tok, hidden=hm, experts_per_rank=e_local, topk=2))

    dist.barrier()
    if rank == 0:
        n_ok = sum(1 for v in res.values() if v.get("ok"))
        print(f"MULTI_SELFTEST {n_ok}/{len(res)} ok")
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"multi_selftest_{world}gpu.json"), "w") as f:
            json.dump(res, f, indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

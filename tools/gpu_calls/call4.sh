#!/bin/bash
# round-2 call 4 (1 GPU): attention bwd with TMA reduce-add dQ, embedding kernels, ncu of both attention kernels, N=1 bench with AdamW CTA budgets
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== selftest"
timeout 600 python tools/gpu_selftest.py attention_train embedding attention_autograd attention_train_perf > gpurun_out/c4_selftest.log 2>&1
echo "rc=$?"; cut -c1-2500 gpurun_out/c4_selftest.log | tail -6
echo "== ncu attention"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention_ -c 4 -f -o gpurun_out/attn_r2c4 python tools/attn_bench.py 8 1024 32 128 0.1 2 > gpurun_out/c4_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/c4_ncu.log
echo "== bench N=1 default"
timeout 300 python bench.py --gpus 1 --steps 8 --warmup 3 --no-e2e > gpurun_out/c4_bench_n1.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c4_bench_n1.log | cut -c1-300
echo "== bench N=1 adamw 148 CTAs"
PFX_ADAMW_CTAS=148 timeout 300 python bench.py --gpus 1 --steps 8 --warmup 3 --no-e2e > gpurun_out/c4_bench_n1_a148.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c4_bench_n1_a148.log | cut -c1-300
echo "== bench N=1 adamw 296 CTAs"
PFX_ADAMW_CTAS=296 timeout 300 python bench.py --gpus 1 --steps 8 --warmup 3 --no-e2e > gpurun_out/c4_bench_n1_a296.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c4_bench_n1_a296.log | cut -c1-300

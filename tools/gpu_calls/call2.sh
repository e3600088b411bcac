#!/bin/bash
# round-2 call 2 (2 GPUs): flash-attention fwd(dropout)/bwd numerics + perf, GEMM regression after the smem headroom change, overlap re-test, benches
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
echo "== selftest (attention + gemm)"
timeout 900 python tools/gpu_selftest.py attention_train attention_autograd attention_fwd gemm_big_sweep gemm_perf_2cta gemm_perf_ffn1 gemm_perf_dgrad gemm_perf_wgrad attention_train_perf > gpurun_out/c2_selftest.log 2>&1
echo "rc=$?"; cut -c1-1800 gpurun_out/c2_selftest.log | tail -12
echo "== nvls selftest (overlap experiment)"
PFX_MULTI_ONLY=nvls timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_selftest.py > gpurun_out/c2_nvls.log 2>&1
echo "rc=$?"; grep -E "overlap|MULTI_SELFTEST|rror" gpurun_out/c2_nvls.log | cut -c1-900 | tail -6
cp gpurun_out/multi_selftest_2gpu.json gpurun_out/c2_multi_selftest_2gpu_nvls.json 2>/dev/null
echo "== bench N=2 own collectives"
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e > gpurun_out/c2_bench_n2_own.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c2_bench_n2_own.log | cut -c1-1600
echo "== bench N=1"
timeout 300 python bench.py --gpus 1 --steps 8 --warmup 3 --no-e2e > gpurun_out/c2_bench_n1.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c2_bench_n1.log | cut -c1-1600
echo "== bench N=1 library attention (A/B)"
PFX_NATIVE_ATTN=0 timeout 300 python bench.py --gpus 1 --steps 8 --warmup 3 --no-e2e > gpurun_out/c2_bench_n1_sdpa.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c2_bench_n1_sdpa.log | cut -c1-1600

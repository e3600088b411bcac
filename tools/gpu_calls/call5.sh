#!/bin/bash
# round-2 call 5 (8 GPUs): headline bench at N=8 through the default path (own NVLS collectives, e2e, named layout child job), NVLS kernels at world 8,
# NCCL baseline at N=8, N=4
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
echo "== bench N=8 default (full contract)"
timeout 900 bash -c "$(declare -f run); run 8 29520 bench.py --gpus 8 --steps 8 --warmup 3" > gpurun_out/c5_bench_n8.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c5_bench_n8.log | cut -c1-3000
echo "== nvls selftest world 8"
PFX_MULTI_ONLY=nvls timeout 400 bash -c "$(declare -f run); run 8 29521 tools/gpu_multi_selftest.py" > gpurun_out/c5_nvls8.log 2>&1
echo "rc=$?"; grep -E "RESULT|MULTI_SELFTEST" gpurun_out/c5_nvls8.log | cut -c1-700 | tail -20
cp gpurun_out/multi_selftest_8gpu.json gpurun_out/c5_multi_selftest_8gpu_nvls.json 2>/dev/null
echo "== bench N=8 NCCL baseline"
timeout 500 bash -c "$(declare -f run); run 8 29522 bench.py --gpus 8 --steps 6 --warmup 3 --no-e2e --p2p 0 --step-overlap 0 --named-layout off" > gpurun_out/c5_bench_n8_nccl.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c5_bench_n8_nccl.log | cut -c1-600
echo "== bench N=4 default"
timeout 500 bash -c "$(declare -f run); run 4 29523 bench.py --gpus 4 --steps 6 --warmup 3 --no-e2e" > gpurun_out/c5_bench_n4.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c5_bench_n4.log | cut -c1-600

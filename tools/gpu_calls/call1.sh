#!/bin/bash
# round-2 call 1 (2 GPUs): NVLS / VMM symmetric memory kernels, overlap experiment, 6.7B bench at N=2 (own collectives vs NCCL), N=1 with / without update overlap
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
echo "== nvls selftest"
PFX_MULTI_ONLY=nvls timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_selftest.py > gpurun_out/c1_nvls.log 2>&1
echo "rc=$?"; grep -E "RESULT|MULTI_SELFTEST|Error|error" gpurun_out/c1_nvls.log | cut -c1-600 | tail -40
cp gpurun_out/multi_selftest_2gpu.json gpurun_out/c1_multi_selftest_2gpu_nvls.json 2>/dev/null
echo "== bench N=2 own collectives"
NCCL_DEBUG=WARN timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 --no-e2e > gpurun_out/c1_bench_n2_own.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/c1_bench_n2_own.log | cut -c1-1500
echo "== bench N=2 NCCL"
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 6 --warmup 3 --no-e2e --p2p 0 --step-overlap 0 > gpurun_out/c1_bench_n2_nccl.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/c1_bench_n2_nccl.log | cut -c1-1500
echo "== bench N=1 overlap on / off"
timeout 300 python bench.py --gpus 1 --steps 6 --warmup 3 --no-e2e > gpurun_out/c1_bench_n1_overlap.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/c1_bench_n1_overlap.log | cut -c1-1500
timeout 300 python bench.py --gpus 1 --steps 6 --warmup 3 --no-e2e --step-overlap 0 > gpurun_out/c1_bench_n1_serial.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/c1_bench_n1_serial.log | cut -c1-1500

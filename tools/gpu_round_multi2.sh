#!/bin/bash
# multi-GPU visit #2: tensor+sequence parallel in the model (NCCL vs fused comm kernels), ZeRO with overlapped all-gather
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m pytest tests/test_gpu_models.py -q -m gpu > gpurun_out/pytest_gpu_models.log 2>&1; echo "pytest models rc=$?"; grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu_models.log | tail -5 | cut -c1-300
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 200)) bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/bench_${name}.log 2>&1; echo "$name rc=$?"; grep '^{' gpurun_out/bench_${name}.log | tail -1 | cut -c1-330; grep -E "Error|error|Traceback" gpurun_out/bench_${name}.log | head -5; }
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_selftest.py > gpurun_out/multi_selftest_$N.log 2>&1; echo "multi selftest rc=$?"
grep -E "RESULT|MULTI_SELFTEST" gpurun_out/multi_selftest_$N.log | grep -E "ag_gemm|zero_p2p|MULTI_SELFTEST|fused_sp" | cut -c1-300
run 6.7b_${N}gpu_sharding_overlap
run 6.7b_${N}gpu_mp2 --layout mp2
run 6.7b_${N}gpu_mp2_fusedtp --layout mp2 --fused-tp 1
if [ "$N" = "8" ]; then run 6.7b_8gpu_mp2_pp2_sharding2 --layout mp2_pp2_sharding2; fi

#!/bin/bash
# 8-GPU visit: headline layouts + the other BASELINE workloads
N=${1:-8}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 200)) "$@"; }
run() { name=$1; shift; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 200)) bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/bench_${name}.log 2>&1; echo "$name rc=$?"; grep '^{' gpurun_out/bench_${name}.log | tail -1 | cut -c1-330; grep -E "Error|Traceback" gpurun_out/bench_${name}.log | head -3; }
wl() { name=$1; shift; timeout 360 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 200)) tools/bench_workloads.py --gpus $N --steps 4 --warmup 2 "$@" > gpurun_out/workload_${name}.log 2>&1; echo "$name rc=$?"; grep '^{' gpurun_out/workload_${name}.log | tail -1 | cut -c1-420; grep -E "Error|Traceback" gpurun_out/workload_${name}.log | sort | uniq -c | head -4; }
run 6.7b_${N}gpu_sharding_overlap
run 6.7b_${N}gpu_mp2_pp2_sharding2 --layout mp2_pp2_sharding2
wl moe_${N}gpu_nccl --workload moe
wl moe_${N}gpu_p2p --workload moe --p2p 1
wl vit_${N}gpu --workload vit
wl ernie_${N}gpu --workload ernie
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_selftest.py > gpurun_out/multi_selftest_$N.log 2>&1; echo "multi selftest rc=$?"
grep -E "RESULT|MULTI_SELFTEST" gpurun_out/multi_selftest_$N.log | grep -E "ag_gemm_perf|gemm_rs_perf|zero_p2p|MULTI_SELFTEST|moe" | cut -c1-330

#!/bin/bash
# round-2 call 17 (2 GPUs): GPT-MoE after the gate's host syncs were removed — 4 layers (compare: 41.1 ms/step in call 13) and the full 24-layer model
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
echo "== moe 2 GPUs, 4 layers"
timeout 300 bash -c "$(declare -f run); run 2 29571 tools/bench_workloads.py --workload moe --gpus 2 --p2p 1 --layers 4 --steps 6 --warmup 3" > gpurun_out/c17_moe_l4.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c17_moe_l4.log | cut -c1-1100
echo "== moe 2 GPUs, full model"
timeout 400 bash -c "$(declare -f run); run 2 29572 tools/bench_workloads.py --workload moe --gpus 2 --p2p 1 --steps 5 --warmup 3" > gpurun_out/c17_moe_full.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c17_moe_full.log | cut -c1-1100

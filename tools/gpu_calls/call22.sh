#!/bin/bash
# round-2 call 22 (1 GPU): GPT-MoE 1.3B on one GPU through the sync-free grouped path (1 -> 8 scaling baseline on the same code path)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 100 python tools/bench_workloads.py --workload moe --gpus 1 --p2p 0 --steps 4 --warmup 3 > gpurun_out/c22_moe_1gpu_loop.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c22_moe_1gpu_loop.log | cut -c1-1000

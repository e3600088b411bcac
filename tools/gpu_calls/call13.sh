#!/bin/bash
# round-2 call 13 (2 GPUs): the other workloads through the updated harness (JSON contract) — GPT-MoE with the sync-free grouped path, ViT ZeRO-2 on own collectives
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
echo "== moe 2 GPUs, 4 layers, grouped sync-free path"
timeout 500 bash -c "$(declare -f run); run 2 29551 tools/bench_workloads.py --workload moe --gpus 2 --p2p 1 --layers 4 --steps 4 --warmup 3" > gpurun_out/c13_moe_p2p.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c13_moe_p2p.log | cut -c1-1500; grep -E "Error|error" gpurun_out/c13_moe_p2p.log | tail -5 | cut -c1-300
echo "== moe 2 GPUs, 4 layers, NCCL all-to-all + expert loop"
timeout 500 bash -c "$(declare -f run); run 2 29552 tools/bench_workloads.py --workload moe --gpus 2 --p2p 0 --layers 4 --steps 4 --warmup 3" > gpurun_out/c13_moe_nccl.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c13_moe_nccl.log | cut -c1-700
echo "== vit 2 GPUs, 4 layers"
timeout 500 bash -c "$(declare -f run); run 2 29553 tools/bench_workloads.py --workload vit --gpus 2 --p2p 1 --layers 4 --steps 4 --warmup 3" > gpurun_out/c13_vit.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c13_vit.log | cut -c1-1200; grep -E "Error|error" gpurun_out/c13_vit.log | tail -5 | cut -c1-300

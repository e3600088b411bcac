#!/bin/bash
# round-2 call 15 (8 GPUs, final): headline bench through the default path incl. the named-layout child job (BASELINE config #2), GPT-MoE 8 x 1.3B on the
# sync-free grouped path, and — only if the lease has time left — ERNIE mp4 x ZeRO-3 with MX fp8 TP GEMMs (config #3)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
T0=$(date +%s)
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
echo "== bench N=8 default (full contract, named layout child)"
PFX_NAMED_LAYOUT_TIMEOUT=200 timeout 420 bash -c "$(declare -f run); run 8 29560 bench.py --gpus 8 --steps 6 --warmup 3" > gpurun_out/c15_bench_n8.log 2>&1
echo "rc=$? t=$(( $(date +%s) - T0 ))s"; tail -1 gpurun_out/c15_bench_n8.log | cut -c1-4000
tail -3 gpurun_out/named_layout_child_rank0.log 2>/dev/null | cut -c1-300
echo "== moe 8 GPUs, sync-free grouped path"
timeout 170 bash -c "$(declare -f run); run 8 29561 tools/bench_workloads.py --workload moe --gpus 8 --p2p 1 --steps 4 --warmup 3" > gpurun_out/c15_moe_p2p.log 2>&1
echo "rc=$? t=$(( $(date +%s) - T0 ))s"; tail -1 gpurun_out/c15_moe_p2p.log | cut -c1-900
if [ $(( $(date +%s) - T0 )) -lt 290 ]; then
  echo "== ernie 8 GPUs mp4 x stage3, MX fp8 TP GEMMs"
  timeout 150 bash -c "$(declare -f run); run 8 29562 tools/bench_workloads.py --workload ernie --gpus 8 --fp8 1 --sp 1 --steps 4 --warmup 3" > gpurun_out/c15_ernie_fp8.log 2>&1
  echo "rc=$? t=$(( $(date +%s) - T0 ))s"; tail -1 gpurun_out/c15_ernie_fp8.log | cut -c1-900
fi

#!/bin/bash
# round-2 call 8 (2 GPUs): grouped expert GEMMs (device-side segment table) unit check + perf, sync-free MoE layer vs NCCL path
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
echo "== selftest"
timeout 600 python tools/gpu_selftest.py moe_grouped moe_grouped_perf gemm_nt_2cta gemm_perf_2cta > gpurun_out/c8_selftest.log 2>&1
echo "rc=$?"; cut -c1-1200 gpurun_out/c8_selftest.log | tail -8
echo "== multi selftest"
PFX_MULTI_ONLY=moe timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/gpu_multi_selftest.py > gpurun_out/c8_multi.log 2>&1
echo "rc=$?"; grep -E "RESULT|MULTI_SELFTEST|rror" gpurun_out/c8_multi.log | cut -c1-1500 | tail -14
cp gpurun_out/multi_selftest_2gpu.json gpurun_out/c8_multi_selftest_2gpu.json 2>/dev/null

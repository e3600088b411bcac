#!/bin/bash
# round-2 call 12b (1 GPU): MX block-scaled fp8 GEMM, staged protocol checks
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/gpu_selftest.py gemm_mxfp8 gemm_fp8 > gpurun_out/c12b_selftest.log 2>&1
echo "rc=$?"; grep -E "pfx|check|Error|error" gpurun_out/c12b_selftest.log | cut -c1-1500 | tail -8

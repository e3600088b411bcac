#!/bin/bash
# round-2 call 9 (1 GPU): grouped GEMM small-shape cases with self-describing argument rejection
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/gpu_selftest.py moe_grouped > gpurun_out/c9_selftest.log 2>&1
echo "rc=$?"; grep -E "pfx gemm|check" gpurun_out/c9_selftest.log | cut -c1-1500 | tail -8

#!/bin/bash
# round-2 call 12 (1 GPU): attention backward for 64-wide heads + L2-friendly CTA order, regression of the 128 path, timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/gpu_selftest.py attention_train_perf attention_autograd > gpurun_out/c12_selftest.log 2>&1
echo "rc=$?"; cut -c1-2500 gpurun_out/c12_selftest.log | tail -4; grep -o '"perf_B[^}]*}' gpurun_out/c12_selftest.log

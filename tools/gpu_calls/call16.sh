#!/bin/bash
# round-2 call 16 (1 GPU): evoformer attention with the native backward kernel, 1-GPU sync-free MoE tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/gpu_selftest.py evoformer_attention > gpurun_out/c16_selftest.log 2>&1
echo "rc=$?"; grep -E "pfx|check|Error|error" gpurun_out/c16_selftest.log | cut -c1-2200 | tail -6
timeout 600 python -m pytest tests/test_gpu_models.py -q -m gpu -k "moe" > gpurun_out/c16_pytest_moe.log 2>&1
echo "rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/c16_pytest_moe.log | tail -8 | cut -c1-400

#!/bin/bash
# round-2 call 23 (1 GPU): MoE GPU tests after the gate launch reduction (expert counts reused, capacity tensor cached)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_models.py -q -m gpu -k "moe" > gpurun_out/c23_pytest_moe.log 2>&1
echo "rc=$?"; grep -E "passed|failed|FAILED|^E " gpurun_out/c23_pytest_moe.log | tail -6 | cut -c1-400

#!/bin/bash
# round-2 call 14b (1 GPU): pytest -m gpu again after the attention routing test was brought up to date (no -x: every failure is listed)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/c14b_pytest_gpu.log 2>&1
echo "rc=$?"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/c14b_pytest_gpu.log | tail -12 | cut -c1-300

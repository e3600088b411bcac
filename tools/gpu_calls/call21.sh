#!/bin/bash
# round-2 call 21 (1 GPU): the evoformer block test under bf16 autocast
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_models.py -q -m gpu -k "evoformer_block" > gpurun_out/c21_pytest.log 2>&1
echo "rc=$?"; grep -E "passed|failed|FAILED|^E " gpurun_out/c21_pytest.log | tail -8 | cut -c1-500

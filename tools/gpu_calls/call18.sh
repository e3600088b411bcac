#!/bin/bash
# round-2 call 18 (1 GPU): where does the host still wait inside a step (torch sync debug mode), evoformer forward+backward timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python tools/gpu_selftest.py sync_audit evoformer_attention_perf > gpurun_out/c18_selftest.log 2>&1
echo "rc=$?"; grep -E "pfx|check|Error|error" gpurun_out/c18_selftest.log | cut -c1-3000 | tail -6

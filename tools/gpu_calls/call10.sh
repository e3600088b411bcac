#!/bin/bash
# round-2 call 10 (1 GPU): grouped GEMM small shapes after the context fix, evoformer gated attention kernel (numerics + perf)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/gpu_selftest.py moe_grouped evoformer_attention evoformer_attention_perf > gpurun_out/c10_selftest.log 2>&1
echo "rc=$?"; grep -E "pfx|check" gpurun_out/c10_selftest.log | cut -c1-2500 | tail -8

#!/bin/bash
# round-2 call 20 (1 GPU): ncu --set full of the kernels added late in the round (MX fp8 GEMM, evoformer backward, grouped expert GEMM)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
cap() { out=$1; kre=$2; skip=$3; shift 3; timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -f -o gpurun_out/$out "$@" > gpurun_out/ncu_$out.log 2>&1; echo "$out rc=$?"; }
cap r2_mxfp8_gemm gemm_lowp_kernel 12 python tools/gpu_selftest.py gemm_mxfp8_perf
cap r2_evoformer_bwd evoformer_attn_bwd_kernel 3 python tools/gpu_selftest.py evoformer_attention_perf
cap r2_grouped_gemm gemm_tcgen05_kernel 30 python tools/gpu_selftest.py moe_grouped_perf
ls -la gpurun_out/r2_*.ncu-rep | awk '{print $5, $9}'

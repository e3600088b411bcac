#!/bin/bash
# ncu --set full captures (one kernel each); reports land in gpurun_out/*.ncu-rep
mkdir -p gpurun_out
cap() { out=$1; kre=$2; skip=$3; shift 3; timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -f -o gpurun_out/$out "$@" > gpurun_out/ncu_$out.log 2>&1; echo "$out rc=$?"; }
cap attention_fwd attention_fwd_kernel 9 python tools/gpu_selftest.py attention_fwd
cap gemm_dual_gelu gemm_tcgen05_kernel 40 python tools/gpu_selftest.py fused_ffn
cap smallm_m16_12288 gemm_smallm_kernel 80 python tools/gpu_selftest.py gemm_smallm
cap attention_decode attention_decode_kernel 30 python tools/gpu_selftest.py attention_decode
cap gemv_w8a8 gemv_w8a8_kernel 10 python tools/gpu_selftest.py gemv_w8a8
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'

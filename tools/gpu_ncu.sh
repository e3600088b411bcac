#!/bin/bash
# ncu --set full captures (one kernel each) of the kernels added/changed this round; reports land in gpurun_out/*.ncu-rep
mkdir -p gpurun_out
cap() { out=$1; kre=$2; skip=$3; shift 3; timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -f -o gpurun_out/$out "$@" > gpurun_out/ncu_$out.log 2>&1; echo "$out rc=$?"; }
cap smallm_m16 gemm_smallm_kernel 9 python tools/gpu_selftest.py gemm_smallm
cap adamw adamw_kernel 3 python tools/gpu_selftest.py adamw
cap norm_bwd norm_bwd_kernel 3 python tools/gpu_selftest.py layernorm
cap gemv gemv_skinny_kernel 3 python tools/gpu_selftest.py gemv_skinny
cap gemm_int8 gemm_lowp_kernel 8 python tools/gpu_selftest.py gemm_int8_perf
cap gemm_wgrad gemm_tcgen05_kernel 8 python tools/gpu_selftest.py gemm_perf_wgrad
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'

#!/bin/bash
# round-2 call 11 (1 GPU): compute-sanitizer evidence (memcheck / racecheck on the new kernels) and ncu captures of the attention kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== memcheck"
timeout 900 bash tools/gpu_sanitize.sh memcheck -- embedding moe_grouped evoformer_attention norm_residual
echo "== racecheck"
timeout 600 bash tools/gpu_sanitize.sh racecheck -- embedding norm_residual
cp gpurun_out/sanitize/summary.txt gpurun_out/sanitize_summary_racecheck.txt 2>/dev/null
echo "== ncu attention"
cap() { out=$1; kre=$2; skip=$3; shift 3; timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -f -o gpurun_out/$out "$@" > gpurun_out/ncu_$out.log 2>&1; echo "$out rc=$?"; }
cap r2_attn_fwd attention_fwd_kernel 1 python tools/attn_bench.py 8 1024 32 128 0.1 2
cap r2_attn_bwd attention_bwd_kernel 1 python tools/attn_bench.py 8 1024 32 128 0.1 2
cap r2_evoformer evoformer_attn_fwd_kernel 2 python tools/gpu_selftest.py evoformer_attention_perf
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_selftest.py attention_fwd 2>&1 | tail -2 | cut -c1-2200
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | tail -8 | cut -c1-300
timeout 600 python tools/gpu_selftest.py gemv_tuning 2>&1 | tail -2 | cut -c1-2500
bash tools/gpu_ncu.sh 2>&1 | tail -8
for chk in gemm_nt_2cta gemm_smallm layernorm adamw; do timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/gpu_selftest.py $chk > gpurun_out/sanitizer_memcheck_$chk.log 2>&1; echo "memcheck $chk rc=$? $(grep -c 'ERROR SUMMARY: 0 errors' gpurun_out/sanitizer_memcheck_$chk.log)"; done
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/gpu_selftest.py layernorm > gpurun_out/sanitizer_racecheck_layernorm.log 2>&1; echo "racecheck layernorm rc=$?"; tail -2 gpurun_out/sanitizer_racecheck_layernorm.log | cut -c1-200

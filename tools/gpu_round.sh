#!/bin/bash
# One GPU-box visit: low-precision GEMM checks, per-kernel launch list of the flagship step, ncu capture of the int8 GEMM.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu_info.csv
timeout 600 python tools/gpu_selftest.py gemm_int8 gemm_int8_pair gemm_fp8 gemm_fp8_pair gemm_int8_perf gemm_fp8_perf 2>&1 | tail -8 | cut -c1-700
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 3200 --csv --log-file gpurun_out/launches_6.7b.csv python bench.py --steps 1 --warmup 1 --no-e2e > gpurun_out/ncu_launches.log 2>&1; echo "launches rc=$?"; wc -l gpurun_out/launches_6.7b.csv

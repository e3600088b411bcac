#!/bin/bash
# One GPU-box visit: tests, small + flagship bench, per-kernel launch list of the flagship step, ncu capture of the GEMM.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu_info.csv
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 400 python tools/gpu_selftest.py gemm_int8 gemm_int8_pair gemm_fp8 gemm_fp8_pair gemm_int8_perf gemm_fp8_perf 2>&1 | tail -8 | cut -c1-600
timeout 600 python bench.py --model gpt-345m --steps 5 --warmup 3 > gpurun_out/bench_345m.log 2>&1; echo "bench345 rc=$?"; tail -2 gpurun_out/bench_345m.log | cut -c1-1500
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_6.7b.log 2>&1; echo "bench6.7 rc=$?"; tail -3 gpurun_out/bench_6.7b.log | cut -c1-1800
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 5600 --csv --log-file gpurun_out/launches_6.7b.csv python bench.py --steps 1 --warmup 1 --no-e2e > gpurun_out/ncu_launches.log 2>&1; echo "launches rc=$?"

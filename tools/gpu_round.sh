#!/bin/bash
# One GPU-box visit: tests, smoke, small + flagship bench, launch list, ncu capture of the GEMM.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu_info.csv
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --model gpt-345m --steps 5 --warmup 3 > gpurun_out/bench_345m.log 2>&1; echo "bench345 rc=$?"; tail -2 gpurun_out/bench_345m.log | cut -c1-1500
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_6.7b.log 2>&1; echo "bench6.7 rc=$?"; tail -3 gpurun_out/bench_6.7b.log | cut -c1-1800
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 8 -c 2 -o gpurun_out/gemm_ffn1 python tools/gpu_selftest.py gemm_perf_ffn1 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv --log-file gpurun_out/launches_345m.csv python bench.py --model gpt-345m --steps 2 --warmup 2 --no-e2e > gpurun_out/ncu_launches.log 2>&1; echo "launches rc=$?"

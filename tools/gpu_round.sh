#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python tools/gpu_selftest.py gemm_smallm gemv_skinny topp 2>&1 | tail -4 | cut -c1-2500
timeout 900 python tools/bench_inference.py --model gpt-6.7b --batches 1,2,4,8,16 --iters 10 > gpurun_out/inference_6.7b.log 2>&1; echo "inference rc=$?"; grep '^{' gpurun_out/inference_6.7b.log | cut -c1-300

#!/bin/bash
# One GPU-box visit: tests, kernel checks, flagship bench, generation latency.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu_info.csv
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_selftest.py layernorm rmsnorm adamw gemv_skinny 2>&1 | tail -6 | cut -c1-1200
timeout 600 python bench.py --model gpt-345m --steps 5 --warmup 3 > gpurun_out/bench_345m.log 2>&1; echo "bench345 rc=$?"; tail -1 gpurun_out/bench_345m.log | cut -c1-400
timeout 1200 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_6.7b.log 2>&1; echo "bench6.7 rc=$?"; tail -1 gpurun_out/bench_6.7b.log | cut -c1-1800
timeout 900 python tools/bench_inference.py --model gpt-6.7b --batches 1,2,4,8,16 --iters 10 > gpurun_out/inference_6.7b.log 2>&1; echo "inference rc=$?"; grep '^{' gpurun_out/inference_6.7b.log | cut -c1-400
timeout 600 python tools/bench_inference.py --model gpt-6.7b --batches 1,8 --iters 10 --no-graph > gpurun_out/inference_6.7b_nograph.log 2>&1; echo "inference-nograph rc=$?"; grep '^{' gpurun_out/inference_6.7b_nograph.log | cut -c1-400
timeout 600 python tools/bench_inference.py --model gpt-345m --batches 1,16 --iters 10 > gpurun_out/inference_345m.log 2>&1; echo "inference345 rc=$?"; grep '^{' gpurun_out/inference_345m.log | cut -c1-400

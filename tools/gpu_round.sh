#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | tail -15 | cut -c1-400
timeout 600 python tools/gpu_selftest.py attention_decode decode_fused topp 2>&1 | tail -4 | cut -c1-1500
timeout 900 python tools/bench_inference.py --model gpt-6.7b --batches 1,2,4,8,16 --iters 10 > gpurun_out/inference_6.7b.log 2>&1; echo "inference rc=$?"; grep '^{' gpurun_out/inference_6.7b.log | cut -c1-230
timeout 900 python tools/bench_inference.py --model gpt-345m --batches 1,2,4,8,16 --iters 10 > gpurun_out/inference_345m.log 2>&1; echo "inference345 rc=$?"; grep '^{' gpurun_out/inference_345m.log | cut -c1-230
timeout 900 python tools/bench_inference.py --model gpt-6.7b --batches 1,16 --iters 10 --int8 > gpurun_out/inference_6.7b_int8.log 2>&1; echo "inference int8 rc=$?"; grep '^{' gpurun_out/inference_6.7b_int8.log | cut -c1-230

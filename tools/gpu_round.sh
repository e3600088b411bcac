#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|^ERROR|Error" gpurun_out/pytest_gpu.log | tail -15 | cut -c1-400
timeout 600 python tools/gpu_selftest.py attention_decode gemv_w8a8 2>&1 | tail -3 | cut -c1-1800
timeout 900 python tools/bench_inference.py --model gpt-6.7b --batches 1,2,4,8,16 --iters 10 > gpurun_out/inference_6.7b.log 2>&1; echo "inference rc=$?"; grep '^{' gpurun_out/inference_6.7b.log | cut -c1-260
timeout 900 python tools/bench_inference.py --model gpt-6.7b --batches 1,2,4,8,16 --iters 10 --int8 > gpurun_out/inference_6.7b_int8.log 2>&1; echo "inference int8 rc=$?"; grep '^{' gpurun_out/inference_6.7b_int8.log | cut -c1-260; grep -E "Error|error" gpurun_out/inference_6.7b_int8.log | head -3
for w in moe vit; do timeout 600 python tools/bench_workloads.py --workload $w --gpus 1 --steps 3 --warmup 2 > gpurun_out/workload_${w}_1gpu.log 2>&1; echo "$w rc=$?"; grep '^{' gpurun_out/workload_${w}_1gpu.log | cut -c1-400; grep -E "Error|error" gpurun_out/workload_${w}_1gpu.log | head -3; done

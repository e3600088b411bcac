#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_selftest.py gemv_skinny 2>&1 | tail -2 | cut -c1-1200
timeout 900 python tools/bench_inference.py --model gpt-6.7b --batches 1,2,4,8,16 --iters 10 > gpurun_out/inference_6.7b.log 2>&1; echo "inference rc=$?"; grep '^{' gpurun_out/inference_6.7b.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 2400 --csv --log-file gpurun_out/launches_decode.csv python tools/bench_inference.py --model gpt-6.7b --batches 1 --iters 1 --warmup 1 --no-graph > gpurun_out/ncu_decode.log 2>&1; echo "decode launches rc=$?"; wc -l gpurun_out/launches_decode.csv

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/bench_inference.py --model gpt-6.7b --batches 1,2,4,8,16 --iters 10 > gpurun_out/inference_6.7b.log 2>&1; echo "inference rc=$?"; grep '^{' gpurun_out/inference_6.7b.log | cut -c1-230
timeout 900 python tools/bench_inference.py --model gpt-6.7b --batches 1,2,4,8,16 --iters 10 --int8 > gpurun_out/inference_6.7b_int8.log 2>&1; echo "inference int8 rc=$?"; grep '^{' gpurun_out/inference_6.7b_int8.log | cut -c1-230
timeout 900 python tools/bench_inference.py --model gpt-345m --batches 1,2,4,8,16 --iters 10 > gpurun_out/inference_345m.log 2>&1; echo "inference345 rc=$?"; grep '^{' gpurun_out/inference_345m.log | cut -c1-230
timeout 1200 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_6.7b.log 2>&1; echo "bench6.7 rc=$?"; tail -1 gpurun_out/bench_6.7b.log | cut -c1-700
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 3300 --csv --log-file gpurun_out/launches_6.7b_v2.csv python bench.py --steps 1 --warmup 1 --no-e2e > gpurun_out/ncu_launches.log 2>&1; echo "launches rc=$?"; wc -l gpurun_out/launches_6.7b_v2.csv

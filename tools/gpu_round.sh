#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | tail -8 | cut -c1-300
timeout 300 python -c "
import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python tools/bench_inference.py --model gpt-6.7b --batches 1,2 --iters 10 > gpurun_out/inference_6.7b.log 2>&1; echo "inference rc=$?"; grep '^{' gpurun_out/inference_6.7b.log | cut -c1-200
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_6.7b.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_6.7b.log | cut -c1-400

#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/gpu_selftest.py fused_ffn 2>&1 | tail -2 | cut -c1-1500
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | tail -8 | cut -c1-300
timeout 1200 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_6.7b.log 2>&1; echo "bench6.7 rc=$?"; tail -1 gpurun_out/bench_6.7b.log | cut -c1-900
PFX_FUSED_FFN=0 timeout 1200 python bench.py --steps 6 --warmup 3 --no-e2e > gpurun_out/bench_6.7b_unfused.log 2>&1; echo "bench6.7 unfused rc=$?"; tail -1 gpurun_out/bench_6.7b_unfused.log | cut -c1-330

#!/bin/bash
# round-2 call 14 (1 GPU): what the driver runs at round end — pytest -m gpu, smoke(), bench.py N=1
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c14_pytest_gpu.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/c14_pytest_gpu.log | cut -c1-400
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/c14_smoke.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/c14_smoke.log | cut -c1-300
echo "== bench N=1"
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/c14_bench_n1.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c14_bench_n1.log | cut -c1-2000
echo "== mx fp8 perf"
timeout 300 python tools/gpu_selftest.py gemm_mxfp8_perf gemm_fp8_perf > gpurun_out/c14_mx_perf.log 2>&1
grep -o "\"perf_8192[^}]*}\|\"tflops[^,]*" gpurun_out/c14_mx_perf.log | head
echo "== reference arm"
timeout 120 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 | tail -1 | cut -c1-400

#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_6.7b_4gpu.log 2>&1; echo "rc=$?"; grep '^{' gpurun_out/bench_6.7b_4gpu.log | tail -1 | cut -c1-400; grep -E "Error|Traceback" gpurun_out/bench_6.7b_4gpu.log | head -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 tools/run_check.py 2>&1 | grep -E "OK|FAIL|run_check" | cut -c1-200

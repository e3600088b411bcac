#!/bin/bash
# round-2 call 19 (2 GPUs): headline bench at N=2 with the final code (unicast symmetric path; the driver's scaling run covers N = 1, 2, 4, 8)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/c19_bench_n2.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c19_bench_n2.log | cut -c1-2200

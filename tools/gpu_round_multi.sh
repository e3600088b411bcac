#!/bin/bash
# multi-GPU visit: peer-memory kernel checks + flagship bench at N GPUs (NCCL ZeRO and P2P ZeRO)
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_selftest.py > gpurun_out/multi_selftest_$N.log 2>&1; echo "multi selftest rc=$?"
grep -E "RESULT|MULTI_SELFTEST|Error|error" gpurun_out/multi_selftest_$N.log | cut -c1-400 | tail -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_6.7b_${N}gpu_nccl.log 2>&1; echo "bench nccl rc=$?"; grep '^{' gpurun_out/bench_6.7b_${N}gpu_nccl.log | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --p2p 1 > gpurun_out/bench_6.7b_${N}gpu_p2p.log 2>&1; echo "bench p2p rc=$?"; grep '^{' gpurun_out/bench_6.7b_${N}gpu_p2p.log | cut -c1-1500; tail -5 gpurun_out/bench_6.7b_${N}gpu_p2p.log | cut -c1-300

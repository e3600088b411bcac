#!/bin/bash
# round-2 call 3 (2 GPUs): kernel timelines of the 6.7B step at N=1 and N=2 (own collectives), CTA-count A/B of the comm kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
echo "== N=1 profile"
timeout 400 python bench.py --gpus 1 --steps 4 --warmup 3 --no-e2e --profile 2 > gpurun_out/c3_bench_n1.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c3_bench_n1.log | cut -c1-400
echo "== N=2 profile"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 --no-e2e --profile 2 > gpurun_out/c3_bench_n2.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c3_bench_n2.log | cut -c1-400
echo "== N=2 RS 296 CTAs"
PFX_RS_CTAS=296 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 6 --warmup 3 --no-e2e > gpurun_out/c3_bench_n2_rs296.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c3_bench_n2_rs296.log | cut -c1-400
echo "== N=2 RS 16 CTAs, bcast 64"
PFX_RS_CTAS=16 PFX_BCAST_CTAS=64 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 6 --warmup 3 --no-e2e > gpurun_out/c3_bench_n2_rs16.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c3_bench_n2_rs16.log | cut -c1-400

#!/bin/bash
# round-2 call 7 (2 GPUs): ping-pong forward attention (P in tensor memory), pipelined backward v2, ZeRO-3 symmetric re-check, 1-GPU bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
echo "== selftest"
timeout 600 python tools/gpu_selftest.py probe_tmem_a attention_fwd attention_train attention_autograd norm_residual embedding attention_train_perf > gpurun_out/c7_selftest.log 2>&1
echo "rc=$?"; cut -c1-400 gpurun_out/c7_selftest.log | tail -12; grep -o '"perf_B8[^}]*}' gpurun_out/c7_selftest.log
echo "== zero3 symm"
PFX_MULTI_ONLY=zero3 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/gpu_multi_selftest.py > gpurun_out/c7_zero3.log 2>&1
echo "rc=$?"; grep -E "RESULT|MULTI_SELFTEST|rror" gpurun_out/c7_zero3.log | cut -c1-420 | tail -8
echo "== bench N=1"
timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/c7_bench_n1.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c7_bench_n1.log | cut -c1-1500
echo "== named-layout child-job mechanism (2 GPUs, 4 layers, forced mp2)"
PFX_NAMED_LAYOUT_FORCE=mp2 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --layers 4 --steps 4 --warmup 3 --no-e2e > gpurun_out/c7_child_mech.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c7_child_mech.log | cut -c1-2500; tail -5 gpurun_out/named_layout_child_rank0.log | cut -c1-300

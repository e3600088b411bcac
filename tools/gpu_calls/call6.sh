#!/bin/bash
# round-2 call 6 (2 GPUs): attention kernels after the instruction diet, TMEM-A probe, full multi-GPU selftest (ZeRO-3 symm, fused TP backward), mp2 fused-TP A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
echo "== selftest"
timeout 600 python tools/gpu_selftest.py probe_tmem_a attention_train attention_fwd embedding attention_train_perf > gpurun_out/c6_selftest.log 2>&1
echo "rc=$?"; cut -c1-300 gpurun_out/c6_selftest.log | tail -8; grep -o '"perf_B8[^}]*}' gpurun_out/c6_selftest.log
echo "== multi selftest (all sections)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_selftest.py > gpurun_out/c6_multi.log 2>&1
echo "rc=$?"; grep -E "RESULT|MULTI_SELFTEST|rror" gpurun_out/c6_multi.log | cut -c1-420 | tail -32
cp gpurun_out/multi_selftest_2gpu.json gpurun_out/c6_multi_selftest_2gpu.json 2>/dev/null
echo "== bench mp2 fused-tp 1"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --layout mp2 --fused-tp 1 --steps 6 --warmup 3 --no-e2e > gpurun_out/c6_bench_mp2_fused.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c6_bench_mp2_fused.log | cut -c1-700
echo "== bench mp2 fused-tp 0"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --layout mp2 --fused-tp 0 --steps 6 --warmup 3 --no-e2e > gpurun_out/c6_bench_mp2_nccl.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/c6_bench_mp2_nccl.log | cut -c1-700

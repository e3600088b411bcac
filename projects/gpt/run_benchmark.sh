#!/bin/bash
# GPT inference latency table: mp8 then single GPU (exported model in ./output)
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=8 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    projects/gpt/benchmark.py --seq_len 128 --iter 10 --mp_degree 8 --model_dir ./output "$@"
python projects/gpt/benchmark.py --seq_len 128 --iter 10 --mp_degree 1 --model_dir ./output "$@"

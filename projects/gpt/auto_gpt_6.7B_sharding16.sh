#!/bin/bash
# GPT 6.7B pre-training, sharding16 (2 nodes x 8 GPUs: run on both with NNODES=2 NODE_RANK=0/1) through tools/auto.py
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=8 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    tools/auto.py -c paddlefleetx_b200/configs/nlp/gpt/auto/pretrain_gpt_6.7B_sharding16.yaml "$@"

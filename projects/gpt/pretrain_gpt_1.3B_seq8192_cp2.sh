#!/bin/bash
# GPT: pretrain_gpt_1.3B_seq8192_cp2 on 8 GPU(s) (context parallelism, beyond the reference)
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=8 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    tools/train.py -c paddlefleetx_b200/configs/nlp/gpt/pretrain_gpt_1.3B_seq8192_cp2.yaml "$@"

#!/bin/bash
# GPT: qat_gpt_345M_mp8 on 8 GPU(s)
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=8 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    tools/train.py -c paddlefleetx_b200/configs/nlp/gpt/qat_gpt_345M_mp8.yaml "$@"

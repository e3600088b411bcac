#!/bin/bash
# GPT generation: generation_gpt_345M_dp8
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=8 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    tasks/gpt/generation.py -c paddlefleetx_b200/configs/nlp/gpt/generation_gpt_345M_dp8.yaml "$@"

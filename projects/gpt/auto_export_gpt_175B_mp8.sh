#!/bin/bash
# GPT 175B: export the mp8 generation model through the auto-parallel entry
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=8 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    tools/auto_export.py -c paddlefleetx_b200/configs/nlp/gpt/auto/generation_gpt_175B_mp8.yaml "$@"

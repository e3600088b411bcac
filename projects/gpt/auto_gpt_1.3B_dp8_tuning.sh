#!/bin/bash
# GPT 1.3B dp8: measured recompute tuning (Tuning.enable) through tools/auto.py
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=8 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    tools/auto.py -c paddlefleetx_b200/configs/nlp/gpt/auto/pretrain_gpt_1.3B_dp8_tuning.yaml "$@"

#!/bin/bash
# GPT 345M: export the quantisation-aware-trained mp2 generation model
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=2 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    tools/auto_export.py -c paddlefleetx_b200/configs/nlp/gpt/auto/qat_generation_gpt_345M_mp2.yaml "$@"

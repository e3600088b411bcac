#!/bin/bash
# GPT: serve the exported tensor-parallel model in ./output on ${MP:-8} GPUs
set -e
cd "$(dirname "$0")/../.."
MP=${MP:-8}
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=${MP} --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    projects/gpt/inference.py --mp_degree ${MP} --model_dir output "$@"

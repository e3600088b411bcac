#!/bin/bash
# ERNIE: serve the exported tensor-parallel model in ./output on 2 GPUs
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=2 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    projects/ernie/inference.py --model_dir ./output --mp_degree 2 "$@"

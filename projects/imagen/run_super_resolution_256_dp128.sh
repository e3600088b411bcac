#!/bin/bash
# Imagen 64 -> 256 super-resolution stage, dp128 (16 nodes x 8 GPUs)
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=8 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    tools/train.py -c paddlefleetx_b200/configs/multimodal/imagen/imagen_super_resolution_256.yaml \
    -o Distributed.dp_degree=128 "$@"

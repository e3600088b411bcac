#!/bin/bash
# Imagen 2B text-to-image 64x64 with the T5-11B text tower, ZeRO-2 sharding8 x dp32 (32 nodes x 8 GPUs)
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=8 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    tools/train.py -c paddlefleetx_b200/configs/multimodal/imagen/imagen_text2im_64x64_T5-11B.yaml \
    -o Distributed.sharding.sharding_stage=2 \
    -o Distributed.dp_degree=32 \
    -o Distributed.sharding.sharding_degree=8 "$@"

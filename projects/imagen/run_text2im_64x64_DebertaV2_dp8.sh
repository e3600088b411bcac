#!/bin/bash
# Imagen text-to-image 64x64 with the DeBERTa-v2 text tower, dp8
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=8 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    tools/train.py -c paddlefleetx_b200/configs/multimodal/imagen/imagen_text2im_64x64_DebertaV2.yaml \
    -o Distributed.dp_degree=8 "$@"

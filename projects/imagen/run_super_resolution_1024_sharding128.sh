#!/bin/bash
# Imagen 256 -> 1024 super-resolution stage, ZeRO-2 sharding8 per node with recompute, fp32
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=8 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    tools/train.py -c paddlefleetx_b200/configs/multimodal/imagen/imagen_super_resolution_1024.yaml \
    -o Distributed.sharding.sharding_stage=2 \
    -o Distributed.sharding.sharding_degree=8 \
    -o Engine.mix_precision.enable=False \
    -o Global.local_batch_size=1 \
    -o Global.micro_batch_size=1 \
    -o Model.use_recompute=True "$@"

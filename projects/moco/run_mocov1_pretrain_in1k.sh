#!/bin/bash
# MoCo v1 pre-training on ImageNet-1k, 8 GPUs
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=8 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    tools/train.py -c paddlefleetx_b200/configs/vis/moco/mocov1_pt_in1k_1n8c.yaml "$@"

#!/bin/bash
# ViT-B/16 384 quantisation-aware fine-tuning, 8 GPUs
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=8 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    tools/train.py -c paddlefleetx_b200/configs/vis/vit/ViT_base_patch16_384_ft_qat_in1k_2n16c_dp_fp16o2.yaml \
    -o Model.model.drop_rate=0.0 \
    -o Data.Train.sampler.batch_size=16 \
    -o Optimizer.lr.learning_rate=5e-05 \
    -o Optimizer.weight_decay=0.0002 "$@"

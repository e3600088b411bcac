#!/bin/bash
# ViT-B/16 384: export the quantisation-aware-trained model
set -e
cd "$(dirname "$0")/../.."
python tools/export.py -c paddlefleetx_b200/configs/vis/vit/ViT_base_patch16_384_ft_qat_in1k_2n16c_dp_fp16o2.yaml \
    -o Model.model.drop_rate=0.0 \
    -o Data.Train.sampler.batch_size=16 \
    -o Optimizer.lr.learning_rate=5e-05 \
    -o Optimizer.weight_decay=0.0002 "$@"

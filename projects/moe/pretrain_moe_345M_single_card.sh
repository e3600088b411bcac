#!/bin/bash
# GPT-MoE: pretrain_moe_345M_single_card on 1 GPU
set -e
cd "$(dirname "$0")/../.."
python tools/train.py -c paddlefleetx_b200/configs/nlp/moe/pretrain_moe_345M_single_card.yaml "$@"

#!/bin/bash
# ERNIE 345M sequence-classification fine-tuning on one GPU
set -e
cd "$(dirname "$0")/../.."
python tools/train.py -c paddlefleetx_b200/configs/nlp/ernie/finetune_ernie_345M_single_card.yaml "$@"

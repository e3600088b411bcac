#!/bin/bash
# ERNIE base (345M recipe) pre-training on one GPU
set -e
cd "$(dirname "$0")/../.."
python tools/train.py -c paddlefleetx_b200/configs/nlp/ernie/pretrain_ernie_base_345M_single_card.yaml "$@"

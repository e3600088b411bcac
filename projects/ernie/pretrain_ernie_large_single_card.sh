#!/bin/bash
# ERNIE: pretrain_ernie_large_single_card on 1 GPU(s)
set -e
cd "$(dirname "$0")/../.."
python tools/train.py -c paddlefleetx_b200/configs/nlp/ernie/pretrain_ernie_large_single_card.yaml "$@"

#!/bin/bash
# ERNIE: finetune_ernie_base on 1 GPU(s)
set -e
cd "$(dirname "$0")/../.."
python tools/train.py -c paddlefleetx_b200/configs/nlp/ernie/finetune_ernie_base.yaml "$@"

#!/bin/bash
# GPT 1.3B pre-training on one GPU through tools/auto.py
set -e
cd "$(dirname "$0")/../.."
python tools/auto.py -c paddlefleetx_b200/configs/nlp/gpt/auto/pretrain_gpt_1.3B_single_card.yaml "$@"

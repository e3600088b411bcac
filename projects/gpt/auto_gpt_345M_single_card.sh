#!/bin/bash
# GPT 345M pre-training on one GPU through tools/auto.py
set -e
cd "$(dirname "$0")/../.."
python tools/auto.py -c paddlefleetx_b200/configs/nlp/gpt/auto/pretrain_gpt_345M_single_card.yaml "$@"

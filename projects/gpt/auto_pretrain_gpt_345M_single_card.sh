#!/bin/bash
# GPT auto-parallel entry: pretrain_gpt_345M_single_card
set -e
cd "$(dirname "$0")/../.."
python tools/auto.py -c paddlefleetx_b200/configs/nlp/gpt/auto/pretrain_gpt_345M_single_card.yaml "$@"

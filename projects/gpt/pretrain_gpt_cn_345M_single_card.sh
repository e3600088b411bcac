#!/bin/bash
# GPT: pretrain_gpt_cn_345M_single_card on 1 GPU(s)
set -e
cd "$(dirname "$0")/../.."
python tools/train.py -c paddlefleetx_b200/configs/nlp/gpt/pretrain_gpt_cn_345M_single_card.yaml "$@"

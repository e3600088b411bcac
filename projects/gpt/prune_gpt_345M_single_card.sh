#!/bin/bash
# GPT: prune_gpt_345M_single_card on 1 GPU(s)
set -e
cd "$(dirname "$0")/../.."
python tools/train.py -c paddlefleetx_b200/configs/nlp/gpt/prune_gpt_345M_single_card.yaml "$@"

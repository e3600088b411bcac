#!/bin/bash
# GPT: finetune_gpt_345M_single_card_glue on 1 GPU(s)
set -e
cd "$(dirname "$0")/../.."
python tools/train.py -c paddlefleetx_b200/configs/nlp/gpt/finetune_gpt_345M_single_card_glue.yaml "$@"

#!/bin/bash
# GPT 345M: offline evaluation of the structurally pruned model
set -e
cd "$(dirname "$0")/../.."
python tools/eval.py -c paddlefleetx_b200/configs/nlp/gpt/eval_pruned_gpt_345M_single_card.yaml "$@"

#!/bin/bash
# GPT 345M offline evaluation (WikiText perplexity / LAMBADA accuracy)
set -e
cd "$(dirname "$0")/../.."
python tools/eval.py -c paddlefleetx_b200/configs/nlp/gpt/eval_gpt_345M_single_card.yaml "$@"

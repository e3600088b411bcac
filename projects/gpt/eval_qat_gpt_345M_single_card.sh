#!/bin/bash
# GPT: eval_qat_gpt_345M_single_card
set -e
cd "$(dirname "$0")/../.."
python tools/eval.py -c paddlefleetx_b200/configs/nlp/gpt/eval_qat_gpt_345M_single_card.yaml "$@"

#!/bin/bash
# GPT: inference_gpt_345M_single_card
set -e
cd "$(dirname "$0")/../.."
python tools/inference.py -c paddlefleetx_b200/configs/nlp/gpt/inference_gpt_345M_single_card.yaml "$@"

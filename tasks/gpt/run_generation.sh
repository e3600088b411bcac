#!/bin/bash
# GPT 345M text generation from a training checkpoint (interactive sampling demo)
set -e
cd "$(dirname "$0")/../.."
python tasks/gpt/generation.py -c paddlefleetx_b200/configs/nlp/gpt/generation_gpt_345M_single_card.yaml "$@"

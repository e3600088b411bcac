#!/bin/bash
# GPT generation: generation_qat_gpt_345M_single_card
set -e
cd "$(dirname "$0")/../.."
python tasks/gpt/generation.py -c paddlefleetx_b200/configs/nlp/gpt/generation_qat_gpt_345M_single_card.yaml "$@"

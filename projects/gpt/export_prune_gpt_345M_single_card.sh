#!/bin/bash
# GPT 345M: export the structurally pruned generation model
set -e
cd "$(dirname "$0")/../.."
python tools/export.py -c paddlefleetx_b200/configs/nlp/gpt/generation_pruned_gpt_345M_single_card.yaml "$@"

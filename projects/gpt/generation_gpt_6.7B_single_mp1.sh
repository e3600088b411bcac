#!/bin/bash
# GPT generation: generation_gpt_6.7B_single_mp1
set -e
cd "$(dirname "$0")/../.."
python tasks/gpt/generation.py -c paddlefleetx_b200/configs/nlp/gpt/generation_gpt_6.7B_single_mp1.yaml "$@"

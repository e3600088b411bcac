#!/bin/bash
# GPT 6.7B: export the generation model on one GPU (auto-parallel entry)
set -e
cd "$(dirname "$0")/../.."
python tools/auto_export.py -c paddlefleetx_b200/configs/nlp/gpt/auto/generation_gpt_6.7B_mp1.yaml "$@"

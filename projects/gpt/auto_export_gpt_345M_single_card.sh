#!/bin/bash
# GPT 345M: export the single-GPU generation model (auto-parallel entry)
set -e
cd "$(dirname "$0")/../.."
python tools/auto_export.py -c paddlefleetx_b200/configs/nlp/gpt/auto/generation_gpt_345M_single_card.yaml "$@"

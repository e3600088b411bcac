#!/bin/bash
# GPT: export_qat_gpt_345M_single_card
set -e
cd "$(dirname "$0")/../.."
python tools/export.py -c paddlefleetx_b200/configs/nlp/gpt/export_qat_gpt_345M_single_card.yaml "$@"

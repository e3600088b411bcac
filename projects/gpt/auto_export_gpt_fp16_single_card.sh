#!/bin/bash
# GPT 345M: 16-bit export of a trained checkpoint (auto-parallel entry)
set -e
cd "$(dirname "$0")/../.."
python tools/auto_export.py -c paddlefleetx_b200/configs/nlp/gpt/auto/export_gpt_fp16_single_card.yaml \
    -o Engine.save_load.output_dir=./serial_model \
    -o Engine.save_load.ckpt_dir=./output/rank_0/model "$@"

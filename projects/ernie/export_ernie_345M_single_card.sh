#!/bin/bash
# ERNIE 345M: export the inference model
set -e
cd "$(dirname "$0")/../.."
python tools/export.py -c paddlefleetx_b200/configs/nlp/ernie/inference_ernie_345M_single_card.yaml "$@"

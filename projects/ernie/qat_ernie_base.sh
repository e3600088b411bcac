#!/bin/bash
# ERNIE: qat_ernie_base on 1 GPU(s)
set -e
cd "$(dirname "$0")/../.."
python tools/train.py -c paddlefleetx_b200/configs/nlp/ernie/qat_ernie_base.yaml "$@"

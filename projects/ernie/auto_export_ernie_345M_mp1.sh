#!/bin/bash
# ERNIE 345M: export through the auto-parallel entry, one GPU
set -e
cd "$(dirname "$0")/../.."
python tools/auto_export.py -c paddlefleetx_b200/configs/nlp/ernie/auto/finetune_ernie_345M_single_card.yaml \
    -o Distributed.mp_degree=1 "$@"

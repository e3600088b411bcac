#!/bin/bash
# ERNIE 345M: export through the auto-parallel entry, tensor parallel over 2 GPUs
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node=2 --master-addr=${MASTER_ADDR:-127.0.0.1} --master-port=${MASTER_PORT:-29500} \
    tools/auto_export.py -c paddlefleetx_b200/configs/nlp/ernie/auto/finetune_ernie_345M_single_card.yaml \
    -o Distributed.mp_degree=2 "$@"

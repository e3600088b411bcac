#!/bin/bash
# GLUE convergence cases: the datasets must be under $DATA_DIR (default ./dataset) and PRETRAINED should name the 345M checkpoint.
set -e
cd "$(dirname "$0")/../../../../../.."
python -c "import torch, yaml, numpy" || { echo "python environment incomplete (torch / yaml / numpy)"; exit 1; }
python -m paddlefleetx_b200.ops.build > /dev/null          # native kernels + C++ index helper, no-op when up to date
test -d "${DATA_DIR:-./dataset}" || { echo "GLUE data not found under ${DATA_DIR:-./dataset}"; exit 1; }

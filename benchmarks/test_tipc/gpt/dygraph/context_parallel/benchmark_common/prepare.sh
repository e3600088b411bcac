#!/bin/bash
# GPT context-parallel cases (beyond the reference matrix): the data ranks share sequences, Ulysses all-to-all or zigzag ring attention.
set -e
cd "$(dirname "$0")/../../../../../.."
python -c "import torch, yaml, numpy" || { echo "python environment incomplete (torch / yaml / numpy)"; exit 1; }
python -m paddlefleetx_b200.ops.build > /dev/null          # native kernels + C++ index helper, no-op when up to date
test -z "$DATA_DIR" || test -d "$DATA_DIR" || { echo "DATA_DIR=$DATA_DIR does not exist (unset it to run on synthetic data)"; exit 1; }

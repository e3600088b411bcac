#!/bin/bash
# ERNIE: serve the exported model in ./output on one GPU
set -e
cd "$(dirname "$0")/../.."
python projects/ernie/inference.py --model_dir ./output --mp_degree 1 "$@"

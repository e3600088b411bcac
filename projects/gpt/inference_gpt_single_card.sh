#!/bin/bash
# GPT: serve the exported model in ./output on one GPU
set -e
cd "$(dirname "$0")/../.."
python projects/gpt/inference.py --mp_degree 1 --model_dir output "$@"

#!/bin/bash
# Imagen 64 -> 256 super-resolution stage on one GPU
set -e
cd "$(dirname "$0")/../.."
python tools/train.py -c paddlefleetx_b200/configs/multimodal/imagen/imagen_super_resolution_256.yaml "$@"

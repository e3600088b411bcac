#!/bin/bash
# Imagen 397M text-to-image 64x64 base model on one GPU
set -e
cd "$(dirname "$0")/../.."
python tools/train.py -c paddlefleetx_b200/configs/multimodal/imagen/imagen_397M_text2im_64x64.yaml "$@"

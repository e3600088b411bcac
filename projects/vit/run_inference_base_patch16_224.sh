#!/bin/bash
# ViT-B/16 224: export a checkpoint placed in ./ckpt (this machine is offline: copy model.pdparams there first), then classify an image
set -e
cd "$(dirname "$0")/../.."
test -d ckpt || { echo "put the ViT-B/16 224 checkpoint under ./ckpt first (no download on an offline machine)"; exit 1; }
echo "step 1: export model"
python tools/export.py -c paddlefleetx_b200/configs/vis/vit/ViT_base_patch16_224_inference.yaml \
    -o Engine.save_load.ckpt_dir=./ckpt/ 
echo "step 2: run ViT inference"
python projects/vit/inference.py --model_dir ./output "$@"

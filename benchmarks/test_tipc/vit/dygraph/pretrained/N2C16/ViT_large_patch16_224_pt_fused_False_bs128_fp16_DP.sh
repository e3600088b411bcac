#!/bin/bash
# ViT-L/16 pretrained on N2C16, data parallel, global batch 128, fused attention False
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=ViT_large_patch16_224_pt_fused_False fp_item=fp16 bs_item=128 run_mode=DP use_fused_attn=False device_num=N2C16
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

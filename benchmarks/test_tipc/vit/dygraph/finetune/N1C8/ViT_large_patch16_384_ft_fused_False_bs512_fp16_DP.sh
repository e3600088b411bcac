#!/bin/bash
# ViT-L/16 finetune on N1C8, data parallel, global batch 512, fused attention False
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=ViT_large_patch16_384_ft_fused_False fp_item=fp16 bs_item=512 run_mode=DP use_fused_attn=False device_num=N1C8
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

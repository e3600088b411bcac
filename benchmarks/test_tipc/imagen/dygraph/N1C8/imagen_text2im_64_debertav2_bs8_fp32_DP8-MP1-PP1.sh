#!/bin/bash
# Imagen text-to-image 64x64 with the DeBERTa-v2 text tower, dp8, fp32
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=imagen_text2im_64_debertav2 fp_item=fp32 dp_degree=8 bs_item=8 run_mode=DP8-MP1-PP1 device_num=N1C8
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

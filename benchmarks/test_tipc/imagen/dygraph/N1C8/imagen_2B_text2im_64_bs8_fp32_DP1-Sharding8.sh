#!/bin/bash
# Imagen 2B text-to-image 64x64, ZeRO-2 over 8 GPUs, fp32
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=imagen_2B_text2im_64 fp_item=fp32 dp_degree=1 sharding_degree=8 sharding_stage=2 bs_item=8 run_mode=DP1-Sharding8 device_num=N1C8
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

#!/bin/bash
# Imagen 64 -> 256 super-resolution, one GPU, fp32
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=imagen_SR256 fp_item=fp32 dp_degree=1 bs_item=1 run_mode=DP1-MP1-PP1 device_num=N1C1
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

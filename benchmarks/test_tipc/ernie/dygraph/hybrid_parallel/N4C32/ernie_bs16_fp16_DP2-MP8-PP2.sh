#!/bin/bash
# ERNIE base on N4C32: dp2 x mp8 x pp2, global batch 16, fp16
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=ernie fp_item=fp16 dp_degree=2 mp_degree=8 pp_degree=2 bs_item=16 micro_bs=2 run_mode=DP2-MP8-PP2 device_num=N4C32
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

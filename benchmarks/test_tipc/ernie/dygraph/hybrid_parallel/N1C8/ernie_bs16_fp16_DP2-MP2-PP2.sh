#!/bin/bash
# ERNIE base on N1C8: dp2 x mp2 x pp2, global batch 16, fp16
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=ernie fp_item=fp16 dp_degree=2 mp_degree=2 pp_degree=2 bs_item=16 micro_bs=2 run_mode=DP2-MP2-PP2 device_num=N1C8
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

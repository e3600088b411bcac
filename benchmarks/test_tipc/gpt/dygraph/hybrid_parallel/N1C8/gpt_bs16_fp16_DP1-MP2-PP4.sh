#!/bin/bash
# GPT (h1024) hybrid parallel on N1C8: dp1 x mp2 x pp4, global batch 16, micro batch 2, fp16
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=gpt fp_item=fp16 dp_degree=1 mp_degree=2 pp_degree=4 bs_item=16 micro_bs=2 run_mode=DP1-MP2-PP4 device_num=N1C8
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

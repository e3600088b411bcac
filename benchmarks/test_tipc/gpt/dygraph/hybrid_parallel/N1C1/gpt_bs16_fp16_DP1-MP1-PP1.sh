#!/bin/bash
# GPT (h1024) hybrid parallel on N1C1: dp1 x mp1 x pp1, global batch 16, micro batch 16, fp16
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=gpt fp_item=fp16 dp_degree=1 mp_degree=1 pp_degree=1 bs_item=16 micro_bs=16 run_mode=DP1-MP1-PP1 device_num=N1C1
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

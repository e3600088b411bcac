#!/bin/bash
# GPT (h1024) hybrid parallel on N1C8: dp8 x mp1 x pp1, global batch 64, micro batch 8, fp32
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=gpt fp_item=fp32 dp_degree=8 mp_degree=1 pp_degree=1 bs_item=64 micro_bs=8 run_mode=DP8-MP1-PP1 device_num=N1C8
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

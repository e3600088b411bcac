#!/bin/bash
# GPT (h1024) hybrid parallel on N4C32: dp4 x mp8 x pp1, global batch 16, micro batch 4, fp16
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=gpt fp_item=fp16 dp_degree=4 mp_degree=8 pp_degree=1 bs_item=16 micro_bs=4 run_mode=DP4-MP8-PP1 device_num=N4C32
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

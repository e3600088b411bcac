#!/bin/bash
# GPT auto-parallel entry on one GPU with recompute, fp32
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=gpt_auto_recompute fp_item=fp32 dp_degree=1 mp_degree=1 pp_degree=1 bs_item=8 micro_bs=8 run_mode=DP1-MP1-PP1 use_recompute=True device_num=N1C1
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

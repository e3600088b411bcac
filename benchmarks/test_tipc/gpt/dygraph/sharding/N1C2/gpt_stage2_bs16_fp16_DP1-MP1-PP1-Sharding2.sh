#!/bin/bash
# GPT ZeRO stage 2 over 2 GPUs, global batch 16, fp16
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=gpt_stage2 fp_item=fp16 dp_degree=1 mp_degree=1 pp_degree=1 sharding_degree=2 sharding_stage=2 bs_item=16 micro_bs=8 run_mode=DP1-MP1-PP1-Sharding2 device_num=N1C2
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

#!/bin/bash
# GPT ZeRO stage 2 over 16 GPUs, global batch 128, fp16
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=gpt_stage2 fp_item=fp16 dp_degree=1 mp_degree=1 pp_degree=1 sharding_degree=16 sharding_stage=2 bs_item=128 micro_bs=8 run_mode=DP1-MP1-PP1-Sharding16 device_num=N2C16
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

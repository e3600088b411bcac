#!/bin/bash
# GPT DP1-MP1-PP1-Sharding8 on N1C8, context parallelism: cp_degree 8 (ring), sequence length 8192
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=gpt_cp_ring fp_item=fp16 dp_degree=1 mp_degree=1 pp_degree=1 sharding_degree=8 sharding_stage=1 cp_degree=8 cp_mode=ring seq_len=8192 bs_item=2 micro_bs=1 run_mode=DP1-MP1-PP1-Sharding8 full_size=1 device_num=N1C8
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

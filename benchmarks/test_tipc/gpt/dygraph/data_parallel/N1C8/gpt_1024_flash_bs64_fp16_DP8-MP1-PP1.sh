#!/bin/bash
# GPT data parallel dp8, sequence length 1024, flash attention True, global batch 64
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=gpt_1024_flash fp_item=fp16 dp_degree=8 mp_degree=1 pp_degree=1 bs_item=64 micro_bs=8 run_mode=DP8-MP1-PP1 seq_len=1024 use_flash_attn=True device_num=N1C8
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

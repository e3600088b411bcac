#!/bin/bash
# GPT DP1-MP8-PP1 on N1C8, sequence parallel False
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=gpt_sp_False fp_item=fp16 dp_degree=1 mp_degree=8 pp_degree=1 bs_item=8 micro_bs=8 run_mode=DP1-MP8-PP1 sequence_parallel=False device_num=N1C8
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

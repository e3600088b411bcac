#!/bin/bash
# GPT DP2-MP8-PP2 on N4C32, sequence parallel False
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=gpt_sp_False fp_item=fp16 dp_degree=2 mp_degree=8 pp_degree=2 bs_item=16 micro_bs=2 run_mode=DP2-MP8-PP2 sequence_parallel=False device_num=N4C32
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

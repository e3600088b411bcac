#!/bin/bash
# GPT-345M fine-tuning on RTE; reports the final acc as the convergence metric
set -e
here="$(cd "$(dirname "$0")" && pwd)"
export model_item=CE_gpt_finetune_RTE fp_item=fp16 bs_item=32 run_mode=DP1-MP1-PP1 task=RTE metric_key=acc device_num=N1C1
bash "$here/../benchmark_common/prepare.sh"
bash "$here/../benchmark_common/run_benchmark.sh" "$@"

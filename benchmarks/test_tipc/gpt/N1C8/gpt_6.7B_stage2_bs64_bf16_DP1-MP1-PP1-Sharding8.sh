#!/bin/bash
# TIPC case (reference benchmarks/test_tipc/gpt/dygraph/**): N1C8/gpt_6.7B_stage2_bs64_bf16_DP1-MP1-PP1-Sharding8
cd "$(dirname "$0")/../../../.."
model_item=gpt_6.7B dp=1 mp=1 pp=1 sharding=8 stage=2 bs=8 max_iter=${max_iter:-50} bash benchmarks/test_tipc/run_benchmark.sh

#!/bin/bash
# TIPC case (reference benchmarks/test_tipc/gpt/dygraph/**): N1C2/gpt_1.3B_stage3_bs16_bf16_DP1-MP1-PP1-Sharding2
cd "$(dirname "$0")/../../../.."
model_item=gpt_1.3B dp=1 mp=1 pp=1 sharding=2 stage=3 bs=8 max_iter=${max_iter:-50} bash benchmarks/test_tipc/run_benchmark.sh

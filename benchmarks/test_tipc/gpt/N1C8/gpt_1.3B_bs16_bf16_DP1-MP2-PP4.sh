#!/bin/bash
# TIPC case (reference benchmarks/test_tipc/gpt/dygraph/**): N1C8/gpt_1.3B_bs16_bf16_DP1-MP2-PP4
cd "$(dirname "$0")/../../../.."
model_item=gpt_1.3B dp=1 mp=2 pp=4 sharding=1 bs=16 micro_bs=2 max_iter=${max_iter:-50} bash benchmarks/test_tipc/run_benchmark.sh

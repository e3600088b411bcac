#!/bin/bash
# TIPC case (reference benchmarks/test_tipc/gpt/dygraph/**): N1C8/gpt_345M_bs64_bf16_DP8-MP1-PP1
cd "$(dirname "$0")/../../../.."
model_item=gpt_345M dp=8 mp=1 pp=1 sharding=1 bs=8 max_iter=${max_iter:-50} bash benchmarks/test_tipc/run_benchmark.sh

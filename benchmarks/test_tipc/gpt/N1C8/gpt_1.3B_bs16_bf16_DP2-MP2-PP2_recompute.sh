#!/bin/bash
# TIPC case (reference benchmarks/test_tipc/gpt/dygraph/**): N1C8/gpt_1.3B_bs16_bf16_DP2-MP2-PP2_recompute
cd "$(dirname "$0")/../../../.."
model_item=gpt_1.3B dp=2 mp=2 pp=2 sharding=1 bs=16 micro_bs=2 use_recompute=True max_iter=${max_iter:-50} bash benchmarks/test_tipc/run_benchmark.sh

#!/bin/bash
# TIPC case (reference benchmarks/test_tipc/gpt/dygraph/**): N1C8/gpt_1.3B_bs8_bf16_DP1-MP8-PP1_sequence_parallel
cd "$(dirname "$0")/../../../.."
model_item=gpt_1.3B dp=1 mp=8 pp=1 sharding=1 bs=8 sequence_parallel=True max_iter=${max_iter:-50} bash benchmarks/test_tipc/run_benchmark.sh

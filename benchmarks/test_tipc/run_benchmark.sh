#!/bin/bash
# TIPC-style throughput/convergence harness (reference benchmarks/test_tipc/*/benchmark_common/run_benchmark.sh): builds a
# tools/train.py command line from environment knobs, runs it under `timeout`, then extracts `ips:` and `loss:` from the log.
#   model_item=gpt_345M dp=8 mp=1 pp=1 sharding=1 stage=1 bs=8 fp=bf16 max_iter=50 bash benchmarks/test_tipc/run_benchmark.sh
set -e
cd "$(dirname "$0")/../.."
model_item=${model_item:-gpt_345M}; dp=${dp:-1}; mp=${mp:-1}; pp=${pp:-1}; sharding=${sharding:-1}; stage=${stage:-1}
bs=${bs:-8}; micro_bs=${micro_bs:-$bs}; max_iter=${max_iter:-50}; use_recompute=${use_recompute:-False}; sp=${sequence_parallel:-False}
case $model_item in
  gpt_345M) cfg=paddlefleetx_b200/configs/nlp/gpt/pretrain_gpt_345M_single_card.yaml;;
  gpt_1.3B) cfg=paddlefleetx_b200/configs/nlp/gpt/pretrain_gpt_1.3B_single_card.yaml;;
  gpt_6.7B) cfg=paddlefleetx_b200/configs/nlp/gpt/pretrain_gpt_6.7B_single_card.yaml;;
  *) echo "unknown model_item $model_item"; exit 2;;
esac
n=$((dp * mp * pp * sharding))
log=${log_file:-./tipc_${model_item}_dp${dp}_mp${mp}_pp${pp}_sh${sharding}.log}
cmd="tools/train.py -c $cfg -o Global.local_batch_size=$bs -o Global.micro_batch_size=$micro_bs -o Engine.max_steps=$max_iter -o Engine.eval_freq=-1 \
 -o Engine.logging_freq=1 -o Engine.save_load.save_steps=-1 -o Distributed.dp_degree=$dp -o Distributed.mp_degree=$mp -o Distributed.pp_degree=$pp \
 -o Distributed.sharding.sharding_degree=$sharding -o Distributed.sharding.sharding_stage=$stage -o Model.use_recompute=$use_recompute \
 -o Model.sequence_parallel=$sp -o Data.Train.dataset.name=SyntheticGPTDataset -o Data.Train.loader.num_workers=0"
if [ "$n" -gt 1 ]; then launcher="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port ${MASTER_PORT:-29533}"; else launcher=python; fi
timeout ${timeout:-15m} $launcher $cmd > "$log" 2>&1 || { tail -20 "$log"; exit 1; }
ips=$(grep -o "ips_total: [0-9]* tokens/s" "$log" | tail -n +5 | awk '{s+=$2; n++} END {if (n) printf "%.0f", s/n}')
loss=$(grep -o "loss: [0-9.]*" "$log" | tail -1 | awk '{print $2}')
echo "model_item=$model_item ngpus=$n ips=${ips:-NA} tokens/s final_loss=${loss:-NA}"

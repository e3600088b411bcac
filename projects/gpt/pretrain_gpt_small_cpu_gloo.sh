#!/bin/bash
# GPT-2 small on CPU over gloo: dp2 x ZeRO-1 sharding2 on 4 processes (BASELINE config #1 — plumbing check that needs no GPU).
set -e
cd "$(dirname "$0")/../.."
python -m torch.distributed.run --nnodes=1 --nproc-per-node=${NPROC:-4} --master-addr=127.0.0.1 --master-port=${MASTER_PORT:-29511} \
    tools/train.py -c paddlefleetx_b200/configs/nlp/gpt/pretrain_gpt_small_synthetic.yaml \
    -o Global.device=cpu -o Engine.mix_precision.enable=False \
    -o Distributed.dp_degree=2 -o Distributed.sharding.sharding_degree=$(( ${NPROC:-4} / 2 )) -o Distributed.sharding.sharding_stage=1 \
    -o Engine.max_steps=${MAX_STEPS:-20} -o Engine.logging_freq=5 -o Engine.eval_freq=-1 -o Engine.save_load.save_steps=-1 "$@"

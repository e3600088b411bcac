#!/bin/bash
# usage: bash run_task.sh CoLA|SST2|MNLI|QNLI|RTE|WNLI|MRPC|QQP|STSB  [extra -o overrides]
task=${1:-SST2}; shift || true
cd "$(dirname "$0")/../../../../.."
python examples/transformer/models/GPT/finetune/run.py -c examples/transformer/models/GPT/finetune/configs/finetune_gpt_345M_single_card_glue.yaml \
  -o Data.Train.dataset.name=$task -o Data.Eval.dataset.name=$task "$@"

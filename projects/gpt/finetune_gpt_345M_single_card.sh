#!/bin/bash
# GPT 345M GLUE fine-tuning on one GPU; the task name selects dataset, metric, loss and class count
set -e
cd "$(dirname "$0")/../.."
TASK=${1:?usage: $0 <CoLA|SST2|MRPC|QQP|STSB|MNLI|QNLI|RTE|WNLI> [MNLI eval split]}
shift
declare -A ROOT_OF=([CoLA]=cola_public [SST2]=SST-2 [MRPC]=MRPC [QQP]=QQP [STSB]=STS-B [MNLI]=multinli_1.0 [QNLI]=QNLI [RTE]=RTE [WNLI]=WNLI)
[ -n "${ROOT_OF[$TASK]}" ] || { echo "Task name not recognized, please input CoLA, SST2, MRPC, QQP, STSB, MNLI, QNLI, RTE, WNLI."; exit 1; }
SPLIT=dev; CLASSES=2; EXTRA=()
case $TASK in
  CoLA) EXTRA=(-o Model.metric.train.name=Mcc -o Model.metric.eval.name=Mcc) ;;
  MRPC) SPLIT=test; EXTRA=(-o Engine.num_train_epochs=5 -o Model.metric.train.name=AccuracyAndF1 -o Model.metric.eval.name=AccuracyAndF1) ;;
  QQP)  EXTRA=(-o Model.metric.train.name=AccuracyAndF1 -o Model.metric.eval.name=AccuracyAndF1) ;;
  STSB) CLASSES=1; EXTRA=(-o Model.metric.train.name=PearsonAndSpearman -o Model.metric.eval.name=PearsonAndSpearman -o Model.loss.train.name=MSELoss -o Model.loss.eval.name=MSELoss) ;;
  MNLI) CLASSES=3; SPLIT=${1:-dev_matched}; [ $# -gt 0 ] && shift ;;
  WNLI) EXTRA=(-o Engine.num_train_epochs=5) ;;
esac
python tools/train.py -c paddlefleetx_b200/configs/nlp/gpt/finetune_gpt_345M_single_card_glue.yaml \
    -o Data.Train.dataset.name=$TASK -o Data.Train.dataset.root=./dataset/${ROOT_OF[$TASK]}/ \
    -o Data.Eval.dataset.name=$TASK -o Data.Eval.dataset.root=./dataset/${ROOT_OF[$TASK]}/ -o Data.Eval.dataset.split=$SPLIT \
    -o Model.num_classes=$CLASSES "${EXTRA[@]}" "$@"

#!/bin/bash
# usage: run_with_retry.sh <gpus> <timeout> <script> <tag>   — retries while gpurun answers busy (exit 3)
gpus=$1; to=$2; script=$3; tag=$4
for attempt in $(seq 1 30); do
  if [ "$gpus" = "1" ]; then gpurun --timeout "$to" -- bash "$script" > "gpurun_out/${tag}.stdout" 2>&1; rc=$?
  else gpurun --gpus "$gpus" --timeout "$to" -- bash "$script" > "gpurun_out/${tag}.stdout" 2>&1; rc=$?; fi
  if [ $rc -ne 3 ]; then echo "attempt $attempt rc=$rc" >> "gpurun_out/${tag}.stdout"; break; fi
  sleep 150
done
echo done >> "gpurun_out/${tag}.stdout"

#!/bin/bash
# compute-sanitizer over the single-GPU kernel checks (one process per tool x check; logs + a one-line verdict each under gpurun_out/sanitize/).
#   tools/gpu_sanitize.sh [racecheck|memcheck|synccheck|initcheck ...] -- [check ...]
# racecheck covers shared-memory hazards between the warp roles of the tcgen05 kernels (TMA producer / MMA issuer / epilogue), memcheck the
# global-memory accesses of the gather / scatter kernels, synccheck the named-barrier and mbarrier usage.  The peer-memory protocols are
# exercised by PFX_DEBUG_POISON=1 (parallel/debug_poison.py) instead: the sanitizer does not follow accesses into another process' memory.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
tools=(); checks=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do tools+=("$1"); shift; done
[ "$1" == "--" ] && shift
checks=("$@")
[ ${#tools[@]} -eq 0 ] && tools=(racecheck memcheck)
[ ${#checks[@]} -eq 0 ] && checks=(gemm_nt_2cta embedding norm_residual attention_train moe_grouped)
mkdir -p gpurun_out/sanitize
for t in "${tools[@]}"; do
  for c in "${checks[@]}"; do
    log=gpurun_out/sanitize/${t}_${c}.log
    timeout 900 compute-sanitizer --tool "$t" --print-limit 20 --report-api-errors no python tools/gpu_selftest.py "$c" > "$log" 2>&1
    rc=$?
    summary=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
    echo "$t $c rc=$rc ${summary:-no summary line}"
  done
done | tee -a gpurun_out/sanitize/summary.txt

#!/usr/bin/env bash
# compute-sanitizer over every loop kernel on small LPs:  gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'
# Logs go to gpurun_out/sanitizer/<tool>_<case>.log, the one-line verdicts to gpurun_out/sanitizer/summary.txt
# (copy the summary into profiles/). Every run has its own timeout: an instrumented cooperative kernel that
# cannot make progress costs one timeout, not the call.
set -u
OUT=gpurun_out/sanitizer
mkdir -p "$OUT"
: > "$OUT/summary.txt"
TOOLS="${SANITIZE_TOOLS:-memcheck racecheck synccheck}"
CASES="${SANITIZE_CASES:-staged persistent fused fused_bounded resident resident_bounded sharded1 sharded2 batched_fast batched_generic}"
T="${SANITIZE_TIMEOUT:-150}"
for tool in $TOOLS; do
  for c in $CASES; do
    log="$OUT/${tool}_${c}.log"
    start=$(date +%s)
    timeout "$T" compute-sanitizer --tool "$tool" --error-exitcode 77 python scripts/sanitize_case.py "$c" > "$log" 2>&1
    rc=$?
    verdict=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
    echo "$tool $c rc=$rc $(( $(date +%s) - start ))s  ${verdict:-no summary line}  | $(grep -E "^(staged|persistent|fused|resident|sharded|batched)" "$log" | tail -1)" | tee -a "$OUT/summary.txt"
  done
done

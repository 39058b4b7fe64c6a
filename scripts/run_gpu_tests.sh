#!/usr/bin/env bash
# Runs every GPU test node in its own process with its own timeout, so that a device-side hang
# costs one timeout instead of the whole call (pytest-timeout cannot interrupt a blocked CUDA call).
# usage: scripts/run_gpu_tests.sh [per-test-timeout-seconds] [pytest -k expression]
T="${1:-150}"
K="${2:-}"
mkdir -p gpurun_out
: > gpurun_out/gpu_tests.log
nodes=$(python -m pytest tests -m gpu --collect-only -q -p no:cacheprovider ${K:+-k "$K"} 2>/dev/null | grep '::')
fail=0
for n in $nodes; do
    start=$(date +%s)
    timeout "$T" python -m pytest "$n" -m gpu -x -q -p no:cacheprovider >> gpurun_out/gpu_tests.log 2>&1
    rc=$?
    echo "$n rc=$rc $(( $(date +%s) - start ))s" | tee -a gpurun_out/gpu_tests.summary
    [ $rc -ne 0 ] && fail=1
done
exit $fail

#!/bin/bash
# compute-sanitizer over the op-level GPU tests (one tool per call: memcheck | racecheck | synccheck | initcheck).
# The N = 8192 attention cases are left out (minutes each under the sanitizer); everything else of tests/test_gpu_ops.py and
# tests/test_gpu_seg_ops.py runs.  Log: gpurun_out/r2_sanitizer_<tool>.log
set -u
TOOL=${1:-memcheck}
mkdir -p gpurun_out
timeout ${2:-900} compute-sanitizer --tool $TOOL --print-limit 20 python -m pytest tests/test_gpu_ops.py tests/test_gpu_seg_ops.py -q -m gpu -k "not 8192" -p no:cacheprovider > gpurun_out/r2_sanitizer_$TOOL.log 2>&1
echo "rc $?"
grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/r2_sanitizer_$TOOL.log | tail -5

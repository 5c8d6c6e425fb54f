#!/bin/bash
# Run every GPU test node in its own process (a trapped kernel poisons the CUDA context of its process only).
# usage: tools/gpu_each.sh <pytest file or node prefix> [more...]   -> gpurun_out/each_<name>.log + summary
mkdir -p gpurun_out
SUMMARY=gpurun_out/each_summary.txt
: > $SUMMARY
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv >> $SUMMARY 2>&1
for target in "$@"; do
  nodes=$(python -m pytest "$target" -m gpu --collect-only -q 2>/dev/null | grep "::")
  for n in $nodes; do
    log=gpurun_out/each_$(echo "$n" | tr '/:[]' '____' | cut -c1-150).log
    timeout 300 python -m pytest "$n" -x -q -m gpu > "$log" 2>&1
    rc=$?
    echo "rc=$rc $n $(grep -E 'assert|Error|error|timeout' "$log" | head -3 | tr '\n' ' ' | cut -c1-300)" >> $SUMMARY
  done
done
cat $SUMMARY

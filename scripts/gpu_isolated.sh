#!/bin/bash
# Run each GPU test node in its own process (a trapped kernel kills the CUDA context of its process only),
# each under its own timeout.  Usage: scripts/gpu_isolated.sh <log> <pytest node ids...>
log=$1; shift
mkdir -p gpurun_out
: > "$log"
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv >> "$log" 2>&1
for node in "$@"; do
  echo "=== $node" >> "$log"
  timeout 300 python -m pytest "$node" -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -40 >> "$log"
  echo "--- exit ${PIPESTATUS[0]}" >> "$log"
done
grep -E "^===|^--- exit|passed|failed|error" "$log" | tail -80

#!/bin/bash
# forward attention with parts switched off (A8_ATTN_DIAG builds), one line per build
mkdir -p gpurun_out
: > gpurun_out/attn_diag.log
for t in "" d1 d2 d3 d4 d7 $EXTRA_TAGS; do
  A8_LIB_TAG=$t A8_BUILD_TAG=$t timeout 120 python scripts/attn_diag.py >> gpurun_out/attn_diag.log 2>&1
done
cat gpurun_out/attn_diag.log

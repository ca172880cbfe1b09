#!/bin/bash
# round 2, run D: full suite + bench after the grouped-launch cache / fp16 gelu' / optimizer fixes
tag=${1:-r2d}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${tag}_tests.log
cp gpurun_out/parity_report.md gpurun_out/${tag}_parity.md 2>/dev/null
timeout 900 python bench.py --no-cpu-baseline --no-incumbent > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -12 gpurun_out/${tag}_tests.log | cut -c1-300; cut -c1-2500 gpurun_out/${tag}_bench.json

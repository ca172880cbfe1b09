#!/bin/bash
# round 2, run P: tap-window positional conv + attention math-loop changes: targeted tests first, then the whole GPU suite,
# the per-kernel table and a short bench
tag=${1:-r2p}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x -k "posconv or attention or attn" > gpurun_out/${tag}_tests_quick.log 2>&1
echo "quick tests exit $?" >> gpurun_out/${tag}_tests_quick.log
tail -15 gpurun_out/${tag}_tests_quick.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${tag}_tests.log
tail -8 gpurun_out/${tag}_tests.log | cut -c1-300
timeout 600 python scripts/kernel_table.py --md gpurun_out/${tag}_kernel_table.md > gpurun_out/${tag}_kernel_table.log 2>&1
grep -i "attention\|posconv\|conv0\|layernorm" gpurun_out/${tag}_kernel_table.md
timeout 600 python bench.py --no-incumbent --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-400 gpurun_out/${tag}_bench.json; echo; tail -3 gpurun_out/${tag}_bench.err | cut -c1-300

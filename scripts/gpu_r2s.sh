#!/bin/bash
# round 2, run S: MMA issue loops with incremental descriptors (plain, window, wgrad window): tests, tap timeline, tables, bench
tag=${1:-r2s}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${tag}_tests.log
tail -6 gpurun_out/${tag}_tests.log | cut -c1-300
timeout 100 python scripts/win_trace.py 2>&1 | head -12
timeout 600 python scripts/kernel_table.py --md gpurun_out/${tag}_kernel_table.md > gpurun_out/${tag}_kernel_table.log 2>&1
grep -i "attention\|posconv\|conv0_\|gemm" gpurun_out/${tag}_kernel_table.md
timeout 600 python scripts/gemm_bench.py > gpurun_out/${tag}_gemm_vs_library.log 2>&1
grep "library" gpurun_out/${tag}_gemm_vs_library.log | cut -c1-160
timeout 600 python bench.py --no-incumbent --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-330 gpurun_out/${tag}_bench.json; echo; grep "gpu ms per step" gpurun_out/${tag}_bench.err | cut -c1-200

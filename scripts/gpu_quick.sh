#!/bin/bash
# quick check: GEMM + kernel tests, bench (no CPU baseline), per-kernel table.  Outputs under gpurun_out/<tag>_*.
tag=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${tag}_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 600 python scripts/kernel_table.py > gpurun_out/${tag}_kernel_table.log 2>&1
tail -3 gpurun_out/${tag}_tests.log; cut -c1-400 gpurun_out/${tag}_bench.json; grep -E "gemm|attention" gpurun_out/${tag}_kernel_table.log

#!/bin/bash
# round 2, run ZA: GELU epilogue with one special-function op per value (erfcx polynomial), LayerNorm forward resident in
# one wave: tests, GEMM table, kernel table, bench, ffn1 trace
tag=${1:-r2za}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${tag}_tests.log
tail -4 gpurun_out/${tag}_tests.log | cut -c1-300
timeout 300 python scripts/gemm_bench.py > gpurun_out/${tag}_gemm_vs_library.log 2>&1
grep "library" gpurun_out/${tag}_gemm_vs_library.log | cut -c1-160
timeout 300 python scripts/kernel_table.py --md gpurun_out/${tag}_kernel_table.md > gpurun_out/${tag}_kernel_table.log 2>&1
grep -i "conv0\|layernorm" gpurun_out/${tag}_kernel_table.md
timeout 100 python scripts/gemm_trace.py ffn1 2>&1 | grep -v "^CTA [123]" | head -14
timeout 300 python bench.py --no-incumbent --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-330 gpurun_out/${tag}_bench.json; echo; grep "gpu ms per step\|e2e ms" gpurun_out/${tag}_bench.err | cut -c1-200

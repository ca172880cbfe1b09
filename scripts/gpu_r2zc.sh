#!/bin/bash
# round 2, run ZC: ffn1 bias gradient from the ffn2 data-gradient GEMM's epilogue (a8_gemm_t::colsum): tests, bench, launch counts
tag=${1:-r2zc}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${tag}_tests.log
tail -4 gpurun_out/${tag}_tests.log | cut -c1-300
timeout 300 python bench.py --no-incumbent --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-330 gpurun_out/${tag}_bench.json; echo; grep "gpu ms per step\|e2e ms" gpurun_out/${tag}_bench.err | cut -c1-200
timeout 300 python bench.py --no-incumbent --no-cpu-baseline --workload ctc > gpurun_out/${tag}_bench_ctc.json 2> gpurun_out/${tag}_bench_ctc.err
cut -c1-330 gpurun_out/${tag}_bench_ctc.json; echo; grep "gpu ms per step\|e2e ms" gpurun_out/${tag}_bench_ctc.err | cut -c1-200

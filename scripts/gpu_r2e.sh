#!/bin/bash
tag=${1:-r2e}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${tag}_tests.log
cp gpurun_out/parity_report.md gpurun_out/${tag}_parity.md 2>/dev/null
timeout 300 python scripts/ctc_bench.py > gpurun_out/${tag}_ctc_bench.log 2>&1
timeout 900 python bench.py --no-cpu-baseline --no-incumbent > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 900 python bench.py --workload ctc --no-cpu-baseline --no-incumbent > gpurun_out/${tag}_bench_ctc.json 2> gpurun_out/${tag}_bench_ctc.err
tail -12 gpurun_out/${tag}_tests.log | cut -c1-300; cat gpurun_out/${tag}_ctc_bench.log; cut -c1-2200 gpurun_out/${tag}_bench.json; cut -c1-400 gpurun_out/${tag}_bench_ctc.json; tail -3 gpurun_out/${tag}_bench_ctc.err | cut -c1-300

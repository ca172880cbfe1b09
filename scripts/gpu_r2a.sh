#!/bin/bash
# round 2, run A: full parity suite (no -x: every failure listed), parity report, GEMM bench with library column, bench lines
tag=${1:-r2a}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${tag}_tests.log
cp gpurun_out/parity_report.md gpurun_out/${tag}_parity.md 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
timeout 600 python scripts/gemm_bench.py > gpurun_out/${tag}_gemm_bench.log 2>&1
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 600 python bench.py --workload ctc --no-cpu-baseline > gpurun_out/${tag}_bench_ctc.json 2> gpurun_out/${tag}_bench_ctc.err
tail -30 gpurun_out/${tag}_tests.log; tail -2 gpurun_out/${tag}_smoke.log; cat gpurun_out/${tag}_gemm_bench.log; cut -c1-3000 gpurun_out/${tag}_bench.json; cut -c1-1500 gpurun_out/${tag}_bench_ctc.json; tail -5 gpurun_out/${tag}_bench_ctc.err

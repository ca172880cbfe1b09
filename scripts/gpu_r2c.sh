#!/bin/bash
# round 2, run C: failing tests again, wgrad timeline traces, ncu of the grouped wgrad kernel
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "optim or grouped or full_size or kernels or tcgen05_gemm" > gpurun_out/${tag}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${tag}_tests.log
cp gpurun_out/parity_report.md gpurun_out/${tag}_parity.md 2>/dev/null
timeout 120 python scripts/gemm_trace.py wgrad > gpurun_out/${tag}_trace_wgrad.log 2>&1
timeout 120 python scripts/gemm_trace.py convw > gpurun_out/${tag}_trace_convw.log 2>&1
timeout 300 python scripts/gemm_bench.py group > gpurun_out/${tag}_group.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:gemm_tc_group -c 1 --csv --page raw --log-file gpurun_out/${tag}_ncu_group.csv python scripts/gemm_bench.py group > gpurun_out/${tag}_ncu.log 2>&1
tail -12 gpurun_out/${tag}_tests.log | cut -c1-300; cat gpurun_out/${tag}_trace_wgrad.log; cat gpurun_out/${tag}_trace_convw.log | head -40; cat gpurun_out/${tag}_group.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_' -c 3 -o gpurun_out/${tag}_attn python scripts/kernel_table.py --once > gpurun_out/${tag}_ncu_attn.log 2>&1
ls -la gpurun_out/ | tail -5

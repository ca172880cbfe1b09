#!/bin/bash
# round 2, run ZF: fused column sums through per-tile shared-memory partials: tests, kernel table rows, bench, launch list
tag=${1:-r2zf}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${tag}_tests.log
tail -3 gpurun_out/${tag}_tests.log | cut -c1-200
timeout 300 python scripts/kernel_table.py --md gpurun_out/${tag}_kernel_table.md > gpurun_out/${tag}_kernel_table.log 2>&1
grep -i "ffn2_dgrad\|attention" gpurun_out/${tag}_kernel_table.md
timeout 300 python bench.py --no-incumbent --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-300 gpurun_out/${tag}_bench.json; echo; grep "gpu ms per step" gpurun_out/${tag}_bench.err | cut -c1-200
timeout 200 python bench.py --ncu-step > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --ncu-step > gpurun_out/${tag}_ncu1.log 2>&1
echo "launch list exit $?"
python scripts/launch_summary.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_summary.md 2>&1; head -14 gpurun_out/${tag}_launch_summary.md

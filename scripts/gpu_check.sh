#!/bin/bash
# Standard GPU check: parity tests, smoke, bench, ncu launch list of one step.  Outputs under gpurun_out/<tag>_*.
tag=${1:-chk}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/${tag}_env.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${tag}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
if [ "$2" = "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${3:-2100} -c ${4:-600} --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
fi
tail -3 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_smoke.log | tail -2; cat gpurun_out/${tag}_bench.json | cut -c1-600

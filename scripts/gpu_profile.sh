#!/bin/bash
# ncu evidence for one warmed-up bench step + the per-kernel table.  Outputs under gpurun_out/<tag>_*.
# Each ncu run is preceded (&&) by the same command without ncu, as B200_PROFILING.md requires.
tag=${1:-prof}
mkdir -p gpurun_out
timeout 600 python scripts/kernel_table.py --md gpurun_out/${tag}_kernel_table.md > gpurun_out/${tag}_kernel_table.log 2>&1
echo "kernel_table exit $?"
# (1) launch list of exactly one step (cudaProfilerStart/Stop inside bench.py --ncu-step)
timeout 300 python bench.py --ncu-step > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --ncu-step > gpurun_out/${tag}_ncu1.log 2>&1
echo "launch list exit $?"
# (2) same step: DRAM traffic and tensor-pipe activity per launch
timeout 900 ncu --profile-from-start off --clock-control none --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
  --log-file gpurun_out/${tag}_traffic.csv python bench.py --ncu-step > gpurun_out/${tag}_ncu2.log 2>&1
echo "traffic exit $?"
# (3) full capture of the tensor-core kernels at the headline shapes (one launch each)
if [ "$2" = "full" ]; then
timeout 300 python scripts/kernel_table.py --once > gpurun_out/${tag}_once.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|attn_' -c 16 \
  -o gpurun_out/${tag}_full python scripts/kernel_table.py --once > gpurun_out/${tag}_ncu3.log 2>&1
echo "full exit $?"
fi
tail -5 gpurun_out/${tag}_kernel_table.log

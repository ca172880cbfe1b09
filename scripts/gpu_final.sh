#!/bin/bash
# end-of-round extras: large-model bench line, --set full capture of the GEMM / attention kernels of the final build
tag=${1:-fin}
mkdir -p gpurun_out
timeout 400 python bench.py --model large --no-cpu-baseline --steps 10 > gpurun_out/${tag}_large.json 2> gpurun_out/${tag}_large.err
echo "large exit $?"; cut -c1-300 gpurun_out/${tag}_large.json
timeout 300 python scripts/kernel_table.py --once > gpurun_out/${tag}_once.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|attn_' -c 16 \
  -o gpurun_out/${tag}_full python scripts/kernel_table.py --once > gpurun_out/${tag}_ncu3.log 2>&1
echo "full exit $?"

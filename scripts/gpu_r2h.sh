#!/bin/bash
# round 2, run H: the evidence set for profiles/ (kernel table, ncu launch list + traffic of one bench step, ncu --set full of
# the tensor-core kernels, GEMM vs library, configs[4] sweep, bench lines).  Every ncu run follows the same command without ncu.
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 900 python bench.py --workload ctc > gpurun_out/${tag}_bench_ctc.json 2> gpurun_out/${tag}_bench_ctc.err
timeout 900 python bench.py --model large --no-cpu-baseline --no-incumbent > gpurun_out/${tag}_bench_large.json 2> gpurun_out/${tag}_bench_large.err
timeout 600 python scripts/kernel_table.py --md gpurun_out/${tag}_kernel_table.md > gpurun_out/${tag}_kernel_table.log 2>&1
timeout 600 python scripts/gemm_bench.py > gpurun_out/${tag}_gemm_vs_library.log 2>&1
timeout 300 python bench.py --ncu-step > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --ncu-step > gpurun_out/${tag}_ncu1.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --profile-from-start off --clock-control none --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --log-file gpurun_out/${tag}_traffic.csv python bench.py --ncu-step > gpurun_out/${tag}_ncu2.log 2>&1
echo "traffic exit $?"
timeout 300 python scripts/kernel_table.py --once > gpurun_out/${tag}_once.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc|attn_|ctc_|conv0_kernel|ln_' -c 40 \
  -o gpurun_out/${tag}_full python scripts/kernel_table.py --once > gpurun_out/${tag}_ncu3.log 2>&1
echo "full exit $?"
timeout 900 python scripts/c5_sweep.py > gpurun_out/${tag}_c5_sweep.md 2> gpurun_out/${tag}_c5_sweep.err
python scripts/launch_summary.py gpurun_out/${tag}_launches.csv | head -40
cut -c1-1500 gpurun_out/${tag}_bench.json; echo; cut -c1-600 gpurun_out/${tag}_bench_large.json; echo; tail -30 gpurun_out/${tag}_kernel_table.log

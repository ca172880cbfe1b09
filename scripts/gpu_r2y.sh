#!/bin/bash
# round 2, run Y: device-side draws (csrc/draws.cu) + GEMM epilogue (aux prefetch, batched staging reads, folded GELU
# constants): full GPU tests, GEMM table vs library, bench in both draw modes, host timeline of the e2e step
tag=${1:-r2y}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${tag}_tests.log
tail -6 gpurun_out/${tag}_tests.log | cut -c1-300
timeout 300 python scripts/gemm_bench.py > gpurun_out/${tag}_gemm_vs_library.log 2>&1
grep "library" gpurun_out/${tag}_gemm_vs_library.log | cut -c1-160
timeout 300 python bench.py --no-incumbent --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-330 gpurun_out/${tag}_bench.json; echo; grep "gpu ms per step\|e2e ms" gpurun_out/${tag}_bench.err | cut -c1-200
A8_DEVICE_DRAWS=1 timeout 300 python bench.py --no-incumbent --no-cpu-baseline > gpurun_out/${tag}_bench_devdraws.json 2> gpurun_out/${tag}_bench_devdraws.err
cut -c1-330 gpurun_out/${tag}_bench_devdraws.json; echo; grep "gpu ms per step\|e2e ms\|host enqueue" gpurun_out/${tag}_bench_devdraws.err | cut -c1-200
timeout 200 python scripts/e2e_bubble.py > gpurun_out/${tag}_e2e_bubble_host.log 2>&1
A8_DEVICE_DRAWS=1 timeout 200 python scripts/e2e_bubble.py > gpurun_out/${tag}_e2e_bubble_dev.log 2>&1
head -30 gpurun_out/${tag}_e2e_bubble_host.log; head -24 gpurun_out/${tag}_e2e_bubble_dev.log

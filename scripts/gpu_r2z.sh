#!/bin/bash
# round 2, run Z: packed fp32 pairs (FFMA2) in the GEMM epilogues and the layer-0 conv, bench defaults to device draws
tag=${1:-r2z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${tag}_tests.log
tail -6 gpurun_out/${tag}_tests.log | cut -c1-300
timeout 300 python scripts/kernel_table.py --md gpurun_out/${tag}_kernel_table.md > gpurun_out/${tag}_kernel_table.log 2>&1
grep -i "conv0\|gemm\|layernorm" gpurun_out/${tag}_kernel_table.md
timeout 300 python bench.py --no-incumbent --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-330 gpurun_out/${tag}_bench.json; echo; grep "gpu ms per step\|e2e ms\|host enqueue" gpurun_out/${tag}_bench.err | cut -c1-200
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2z_bench.json"))
print({k: d[k] for k in ("ms_per_step", "value", "host_enqueue_ms_per_step", "e2e_host_draws")}, d["e2e"], d["roofline"]["frac"])
PY

#!/bin/bash
# N-GPU run: NCCL tests (N=2 only), bench at N GPUs with the arena wrapper; NCCL algorithm lines
tag=${1:-r2j}; n=${2:-2}
mkdir -p gpurun_out
if [ "$n" = "2" ]; then
timeout 900 python -m pytest tests/test_ddp.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${tag}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py > gpurun_out/${tag}_dp_check.log 2>&1
fi
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=COLL,TUNING timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $n --steps 20 --warmup 3 --workload ctc > gpurun_out/${tag}_bench_ctc_n${n}.json 2> gpurun_out/${tag}_bench_ctc_n${n}.err
tail -4 gpurun_out/${tag}_tests.log 2>/dev/null | cut -c1-200; tail -2 gpurun_out/${tag}_dp_check.log 2>/dev/null | cut -c1-200
cut -c1-500 gpurun_out/${tag}_bench_n${n}.json; echo; cut -c1-400 gpurun_out/${tag}_bench_ctc_n${n}.json; echo
grep "AllReduce:.*Bytes" gpurun_out/${tag}_bench_n${n}.err | sort | uniq -c | sort -rn | head -5 | cut -c1-200
grep "gpu ms per step" gpurun_out/${tag}_bench_n${n}.err | cut -c1-200

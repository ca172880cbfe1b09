#!/bin/bash
# round 2, run F (2 GPUs): NCCL data-parallel tests, dp_check vs DDP, bench at N=2 (arena) and N=2 (stock DDP)
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddp.py tests/test_ctc.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${tag}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py > gpurun_out/${tag}_dp_check.log 2>&1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
A8_DP=ddp timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n2_ddp.json 2> gpurun_out/${tag}_bench_n2_ddp.err
timeout 600 python scripts/ctc_bench.py > gpurun_out/${tag}_ctc_bench.log 2>&1
tail -6 gpurun_out/${tag}_tests.log | cut -c1-300; tail -4 gpurun_out/${tag}_dp_check.log | cut -c1-300; cut -c1-700 gpurun_out/${tag}_bench_n2.json; echo; cut -c1-400 gpurun_out/${tag}_bench_n2_ddp.json; echo; grep -i "algo\|proto\|channel" gpurun_out/${tag}_bench_n2.err | head -12 | cut -c1-250; tail -3 gpurun_out/${tag}_bench_n2.err | cut -c1-300; cat gpurun_out/${tag}_ctc_bench.log

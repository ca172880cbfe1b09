#!/bin/bash
# N-GPU bench under a few NCCL settings (algorithm / channel count): which all-reduce hides best under backward
n=${1:-8}; tag=${2:-r02}
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=COLL timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n \
    --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $n --steps 20 --warmup 3 --no-incumbent \
    > gpurun_out/${tag}_nccl_${name}_n${n}.json 2> gpurun_out/${tag}_nccl_${name}_n${n}.err
  ms=$(python -c "import json,sys; print(json.loads(open('gpurun_out/${tag}_nccl_${name}_n${n}.json').read().strip().splitlines()[-1])['ms_per_step'])" 2>/dev/null)
  algo=$(grep "AllReduce: 36" gpurun_out/${tag}_nccl_${name}_n${n}.err | head -1 | sed 's/.*-> //' | cut -c1-80)
  echo "$name: ms_per_step=$ms  big all-reduce: $algo" | tee -a gpurun_out/${tag}_nccl_sweep_n${n}.txt
  grep -v "NCCL INFO" gpurun_out/${tag}_nccl_${name}_n${n}.err | tail -3 > gpurun_out/${tag}_nccl_${name}_n${n}.tail; rm gpurun_out/${tag}_nccl_${name}_n${n}.err
}
rm -f gpurun_out/${tag}_nccl_sweep_n${n}.txt
run default A8_X=0
run nvls NCCL_ALGO=NVLS
run nvls16 NCCL_ALGO=NVLS NCCL_MAX_NCHANNELS=16
run ring16 NCCL_MAX_NCHANNELS=16
run ring8 NCCL_MAX_NCHANNELS=8
run nvls_cta8 NCCL_ALGO=NVLS NCCL_MAX_CTAS=8

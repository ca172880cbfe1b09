#!/bin/bash
# N-GPU evidence: bench.py at N GPUs (pretrain, and CTC / large when asked), NCCL-only comparison, NCCL tests at N=2
n=${1:-2}; tag=${2:-r02}; extra=${3:-}
mkdir -p gpurun_out
if [ "$n" = "2" ]; then
  timeout 600 python -m pytest tests/test_ddp.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${tag}_ddp_tests.log 2>&1
  tail -2 gpurun_out/${tag}_ddp_tests.log | cut -c1-200
fi
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $n --steps 20 --warmup 3 --no-incumbent "$@" > gpurun_out/${tag}_${name}_n${n}.json 2> gpurun_out/${tag}_${name}_n${n}.err
  echo "$name: $(cut -c1-220 gpurun_out/${tag}_${name}_n${n}.json)"
  grep "gpu ms per step" gpurun_out/${tag}_${name}_n${n}.err | cut -c1-160
}
run bench A8_X=0 --
if [[ "$extra" == *nccl* ]]; then run bench_nccl A8_ALLREDUCE=nccl --; fi
if [[ "$extra" == *ctc* ]]; then run bench_ctc A8_X=0 -- --workload ctc; fi
if [[ "$extra" == *large* ]]; then run bench_large A8_X=0 -- --model large --no-cpu-baseline; fi
if [[ "$extra" == *ddp* ]]; then run bench_ddp A8_DP=ddp --; fi

#!/bin/bash
# multi-GPU bench through torchrun (one rank per GPU): usage bash tools/gpu_multi.sh TAG N [extra bench args]
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "exit $?"; tail -3 gpurun_out/${TAG}_bench_n$N.err; cat gpurun_out/${TAG}_bench_n$N.json

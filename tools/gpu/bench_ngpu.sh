#!/bin/bash
# bench.py on N GPUs of one box (strong scaling on BASELINE configs[2] by default); extra arguments go to bench.py
mkdir -p gpurun_out
N=${1:-8}; shift
TAG=${TAG:-n$N}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-cpu-baseline --steps 2 --warmup 3 "$@" > gpurun_out/bench_$TAG.log 2>&1; echo "$TAG rc=$?"; tail -1 gpurun_out/bench_$TAG.log | cut -c1-600

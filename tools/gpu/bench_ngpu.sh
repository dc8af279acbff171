#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.log 2>&1; echo "n$N rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-400

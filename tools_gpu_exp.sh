#!/bin/bash
mkdir -p gpurun_out
for R in 0 1 0 1; do
SDD_CONV_RAW=$R timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3 > gpurun_out/bench_ab_$R.log 2>&1; echo "RAW=$R rc=$?"; tail -1 gpurun_out/bench_ab_$R.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['clocks'])"
done

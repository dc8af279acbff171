#!/bin/bash
# same-box A/B of the whole sampler: baseline library (tools/_lib_base.so, built from HEAD) vs the working tree's library
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
A="--steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline"
for i in 1 2; do
  SDD_LIB=$PWD/tools/_lib_base.so timeout 300 python bench.py $A > gpurun_out/ab_bench_base_$i.json 2>gpurun_out/err.log; python -c "import json; d=json.load(open('gpurun_out/ab_bench_base_$i.json')); print('base', d['value'], d['clocks'])"
  timeout 300 python bench.py $A > gpurun_out/ab_bench_new_$i.json 2>gpurun_out/err.log; python -c "import json; d=json.load(open('gpurun_out/ab_bench_new_$i.json')); print('new ', d['value'], d['clocks'])"
done

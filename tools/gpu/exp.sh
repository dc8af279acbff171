#!/bin/bash
# programmatic dependent launch on the forward chain: tests, then same-box A/B against the -DSDD_PDL=0 build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
A="--steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline"
for i in 1 2; do
  SDD_LIB=$PWD/tools/_lib_base.so timeout 300 python bench.py $A > gpurun_out/ab_bench_base_$i.json 2>gpurun_out/err.log; python -c "import json; d=json.load(open('gpurun_out/ab_bench_base_$i.json')); print('nopdl', d['value'], d['clocks'])"
  timeout 300 python bench.py $A > gpurun_out/ab_bench_new_$i.json 2>gpurun_out/err.log; python -c "import json; d=json.load(open('gpurun_out/ab_bench_new_$i.json')); print('pdl  ', d['value'], d['clocks'])"
done
for i in 1 2; do
SDD_LIB=$PWD/tools/_lib_base.so timeout 300 python bench.py $A --batch 8 > gpurun_out/ab_bench_base_b8_$i.json 2>gpurun_out/err.log; python -c "import json; d=json.load(open('gpurun_out/ab_bench_base_b8_$i.json')); print('nopdl b8', d['value'])"
timeout 300 python bench.py $A --batch 8 > gpurun_out/ab_bench_new_b8_$i.json 2>gpurun_out/err.log; python -c "import json; d=json.load(open('gpurun_out/ab_bench_new_b8_$i.json')); print('pdl   b8', d['value'])"
done

#!/bin/bash
# same-box A/B: 4 TMEM accumulators (working tree) vs 2 (-DSDD_CONV_ACCS=2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "conv3x3 or unet_forward or k5 or bench_shape" > gpurun_out/pytest_exp.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_exp.log | cut -c1-300
A2=$PWD/tools/_lib_acc2.so
for i in 1 2; do
  python tools/conv_layers.py > gpurun_out/ab_acc4_$i.log 2>&1; tail -5 gpurun_out/ab_acc4_$i.log
  SDD_LIB=$A2 python tools/conv_layers.py > gpurun_out/ab_acc2_$i.log 2>&1; tail -5 gpurun_out/ab_acc2_$i.log
done
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_acc4.log 2>&1; head -c 250 gpurun_out/bench_acc4.log; echo
SDD_LIB=$A2 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_acc2.log 2>&1; head -c 250 gpurun_out/bench_acc2.log; echo

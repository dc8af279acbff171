#!/bin/bash
# same-box A/B: previous commit's conv kernel (branchy border path, runtime contig) vs the working tree
mkdir -p gpurun_out
PREV=$PWD/tools/_lib_prev.so
for i in 1 2; do
  python tools/conv_layers.py > gpurun_out/ab_new_$i.log 2>&1; tail -5 gpurun_out/ab_new_$i.log
  SDD_LIB=$PREV python tools/conv_layers.py > gpurun_out/ab_prev_$i.log 2>&1; tail -5 gpurun_out/ab_prev_$i.log
done
RES=128 CHUNK=16 python tools/conv_layers.py > gpurun_out/ab_new_128.log 2>&1; tail -5 gpurun_out/ab_new_128.log
RES=128 CHUNK=16 SDD_LIB=$PREV python tools/conv_layers.py > gpurun_out/ab_prev_128.log 2>&1; tail -5 gpurun_out/ab_prev_128.log

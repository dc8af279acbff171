#!/bin/bash
# same-box A/B: baseline library (tools/_lib_base.so, built from HEAD) vs the working tree's library
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "conv or unet_forward or border or nonsquare or non_square" 2>&1 | tail -3
for i in 1 2; do
  SDD_LIB=$PWD/tools/_lib_base.so timeout 200 python tools/conv_layers.py > gpurun_out/ab_base_$i.txt 2>&1; cat gpurun_out/ab_base_$i.txt
  timeout 200 python tools/conv_layers.py > gpurun_out/ab_new_$i.txt 2>&1; cat gpurun_out/ab_new_$i.txt
done

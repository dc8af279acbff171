#!/bin/bash
# same-box A/B: working-tree library vs the -DSDD_CONV_XFORM_FIRST=1 build (tools/_lib_xf.so)
mkdir -p gpurun_out
SDD_LIB=$PWD/tools/_lib_xf.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "conv or unet_forward or border or nonsquare or non_square" 2>&1 | tail -2
for i in 1 2; do
  timeout 200 python tools/conv_layers.py > gpurun_out/ab_base_$i.txt 2>&1; grep -v "^lib" gpurun_out/ab_base_$i.txt
  SDD_LIB=$PWD/tools/_lib_xf.so timeout 200 python tools/conv_layers.py > gpurun_out/ab_new_$i.txt 2>&1; grep -v "^lib" gpurun_out/ab_new_$i.txt
done

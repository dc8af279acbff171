#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -x -k "conv3x3_fused or unet_forward or k3" > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; grep -v "timed out" gpurun_out/pytest_gpu.log | tail -3
if [ $rc -ne 0 ]; then exit 1; fi
TRACE=0 timeout 300 python tools/conv_exp.py > gpurun_out/conv_exp.log 2>&1; echo "exp rc=$?"; grep -v "timed out" gpurun_out/conv_exp.log

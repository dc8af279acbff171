#!/bin/bash
mkdir -p gpurun_out
SDD_CONV_V4=1 timeout 600 python -m pytest tests -m gpu -q --timeout 120 -x -k "conv3x3_fused or unet_forward or k3" > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; grep -v "timed out" gpurun_out/pytest_gpu.log | tail -4
if [ $rc -ne 0 ]; then grep "timed out" gpurun_out/pytest_gpu.log | awk '{print $7, $9, $11, $13}' | awk '{print int($2/32), $3, $4}' | sort | uniq -c | head; exit 1; fi
SDD_CONV_V4=1 TRACE=0 timeout 300 python tools/conv_exp.py > gpurun_out/conv_exp_v4.log 2>&1; echo "exp rc=$?"; grep -v "timed out" gpurun_out/conv_exp_v4.log

#!/bin/bash
mkdir -p gpurun_out
SDD_CONV_RAW=2 timeout 600 python -m pytest tests -m gpu -q --timeout 120 -x -k "conv3x3_fused or unet_forward or k3 or k5" > gpurun_out/pytest_raw.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|assert|timed out|sdd:" gpurun_out/pytest_raw.log | tail -10
for R in 0 2; do
echo "SDD_CONV_RAW=$R"
SDD_CONV_RAW=$R TRACE=0 timeout 300 python tools/conv_exp.py 2>&1 | grep -v "timed out" | cut -c1-260
done

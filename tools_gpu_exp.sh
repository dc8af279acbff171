#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -x -k "conv3x3 or unet_forward or k3 or k5 or shard" > gpurun_out/pytest_raw.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_raw.log
QUICK=1 TRACE=0 timeout 300 python tools/conv_exp.py 2>&1 | grep -v "timed out" | cut -c1-160
CHUNK=64 QUICK=1 TRACE=0 timeout 300 python tools/conv_exp.py 2>&1 | grep -v "timed out" | cut -c1-160

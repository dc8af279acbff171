#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -x -k "conv3x3 or unet_forward or k3 or k5 or shard or chunking" > gpurun_out/pytest_raw.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_raw.log
for V in 0 1 0 1; do
echo "table=$V"
if [ $V = 0 ]; then export SDD_CONV_NO_AB=1; else unset SDD_CONV_NO_AB; fi
CHUNK=64 QUICK=1 TRACE=0 timeout 300 python tools/conv_exp.py 2>&1 | grep -v "timed out" | grep -- "->" | cut -c1-40
done

#!/bin/bash
mkdir -p gpurun_out
cp tools/_libB.so super-diff-disease_b200/libsdd_b200.so
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -x -k "conv3x3 or unet_forward or k3 or k5 or shard or chunking or ragged" > gpurun_out/pytest_raw.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_raw.log
for V in A B A B; do
cp tools/_lib$V.so super-diff-disease_b200/libsdd_b200.so
echo "variant $V"
CHUNK=64 QUICK=1 TRACE=0 timeout 300 python tools/conv_exp.py 2>&1 | grep -v "timed out" | grep -- "->" | cut -c1-42
done

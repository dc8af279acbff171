#!/bin/bash
mkdir -p gpurun_out
for V in A B A B; do
cp tools/_lib$V.so super-diff-disease_b200/libsdd_b200.so
echo "variant $V"
CHUNK=64 QUICK=1 TRACE=0 timeout 300 python tools/conv_exp.py 2>&1 | grep -v "timed out" | grep -- "->" | cut -c1-60
done

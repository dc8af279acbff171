#!/bin/bash
mkdir -p gpurun_out
for V in 0 d 0 d; do
echo "contig=$V"
if [ $V = 0 ]; then export SDD_CONV_CONTIG=0; else unset SDD_CONV_CONTIG; fi
CHUNK=64 QUICK=1 TRACE=0 timeout 300 python tools/conv_exp.py 2>&1 | grep -v "timed out" | grep -- "64->" | cut -c1-40
done

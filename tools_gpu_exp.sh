#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; grep -v "timed out" gpurun_out/pytest_gpu.log | tail -5
if [ $rc -ne 0 ]; then grep "timed out" gpurun_out/pytest_gpu.log | awk '{print $7, $9, $11, $13}' | awk '{print int($2/32), $3, $4}' | sort | uniq -c | head; exit 1; fi
timeout 300 python tools/conv_exp.py > gpurun_out/conv_exp.log 2>&1; echo "exp rc=$?"; grep -v "timed out" gpurun_out/conv_exp.log
SDD_CONV_PREFETCH=0 TRACE=0 timeout 300 python tools/conv_exp.py > gpurun_out/conv_exp_p0.log 2>&1; echo "exp rc=$?"; grep -v "timed out" gpurun_out/conv_exp_p0.log

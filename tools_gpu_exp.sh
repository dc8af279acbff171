#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; grep -v "timed out" gpurun_out/pytest_gpu.log | tail -6
grep superposed_fullres gpurun_out/parity_report.jsonl | tail -2

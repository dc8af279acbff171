#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 60 -x -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?"; grep -E "PARITY.*attention|passed|failed|Error|assert|timed out|sdd:" gpurun_out/pytest_attn.log | tail -30
timeout 120 python tools/attn_bench.py 2>&1 | tail -5

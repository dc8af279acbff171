#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -x -k "n4 or and_" > gpurun_out/pytest_n4.log 2>&1; echo "pytest rc=$?"; grep -E "PARITY.*n4|passed|failed|Error|assert" gpurun_out/pytest_n4.log | tail -30

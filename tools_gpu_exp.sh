#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -x -k "and_ or update or k3 or k5 or graph" > gpurun_out/pytest_and.log 2>&1; echo "pytest rc=$?"; grep -E "PARITY.*and_|passed|failed|Error|assert" gpurun_out/pytest_and.log | tail -30

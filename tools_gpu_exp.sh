#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -x -k "and_ or cli" > gpurun_out/pytest_and.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_and.log; grep "and_three" gpurun_out/pytest_and.log | head -2

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -x -k "ragged" > gpurun_out/pytest_rag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_rag.log; grep "PARITY.*ragged" gpurun_out/pytest_rag.log | cut -c1-220

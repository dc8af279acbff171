#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q --timeout 60 -x -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?"; grep -E "PARITY.*attention_block|passed|failed|Error|assert |timed out|sdd:" gpurun_out/pytest_attn.log | cut -c1-220 | tail -12

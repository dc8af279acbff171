#!/bin/bash
# first-light GPU script: tests, short bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 1 --warmup 1 --diffusion-steps 10 --no-cpu-baseline > gpurun_out/bench_short.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/bench_short.log

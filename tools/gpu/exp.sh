#!/bin/bash
# experiment of the moment: deferred finalisation of the update step -- full suite, sweep, bench
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl gpurun_out/parity_growth.json
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python tools/update_sweep.py > gpurun_out/update_sweep.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/update_sweep.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_exp.log 2>&1; head -c 250 gpurun_out/bench_exp.log; echo

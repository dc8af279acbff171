#!/bin/bash
# experiment of the moment: full GPU suite after the forked-branch change, small-batch / other-config bench lines
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl gpurun_out/parity_growth.json
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --batch 8 > gpurun_out/bench_b8.log 2>&1; head -c 250 gpurun_out/bench_b8.log; echo
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --res 128 --diffusion-steps 100 --batch 16 > gpurun_out/bench_c2.log 2>&1; head -c 250 gpurun_out/bench_c2.log; echo
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --res 512 --diffusion-steps 1000 --batch 4 > gpurun_out/bench_c4.log 2>&1; head -c 250 gpurun_out/bench_c4.log; echo

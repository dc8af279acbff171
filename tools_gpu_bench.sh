#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_full.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log

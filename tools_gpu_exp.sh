#!/bin/bash
mkdir -p gpurun_out
RED="--batch 16 --diffusion-steps 3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-roofline"
timeout 600 python bench.py $RED > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py $RED > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"

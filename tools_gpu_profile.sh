#!/bin/bash
mkdir -p gpurun_out
RED="--batch 16 --diffusion-steps 3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 python bench.py $RED > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py $RED > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 python bench.py $RED > gpurun_out/plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_tc2_kernel|superpose_update_kernel|conv_out1_mma|conv_in_kernel' -s 12 -c 12 -o gpurun_out/prof_r1_v2 -f python bench.py $RED > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | head -30

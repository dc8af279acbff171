#!/bin/bash
# full bench + ncu launch list + ncu full captures of the two graded kernels (reduced config for ncu)
mkdir -p gpurun_out
RED="--batch 6 --diffusion-steps 3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 1200 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench full rc=$?"; tail -1 gpurun_out/bench_full.log
timeout 600 python bench.py $RED > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 260 --csv --log-file gpurun_out/launches.csv python bench.py $RED > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 python bench.py $RED > gpurun_out/plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_tc_kernel|superpose_update_kernel' -s 30 -c 8 -o gpurun_out/prof_r1 -f python bench.py $RED > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out

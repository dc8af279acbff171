#!/bin/bash
mkdir -p gpurun_out
RED="--batch 16 --diffusion-steps 3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 python bench.py $RED > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py $RED > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
CHUNK=64 ITERS=2 timeout 300 python tools/conv_one.py > gpurun_out/one.log 2>&1 && cat gpurun_out/one.log && \
CHUNK=64 ITERS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_tc4_kernel' -s 2 -c 1 -o gpurun_out/prof_conv4_128 -f python tools/conv_one.py > gpurun_out/ncu_one.log 2>&1; echo "ncu rc=$?"
B=64 timeout 300 python tools/update_one.py > gpurun_out/upd_one.log 2>&1 && cat gpurun_out/upd_one.log && \
B=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'superpose_update_kernel' -s 4 -c 1 -o gpurun_out/prof_update_v4 -f python tools/update_one.py > gpurun_out/ncu_upd.log 2>&1; echo "ncu rc=$?"

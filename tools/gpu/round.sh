#!/bin/bash
# Round-end evidence run (1 GPU): full GPU test suite, bench (own arm + reference arm), c5 sweep, ncu launch list,
# ncu --set full of the two roofline kernels.  Every ncu pass runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_full.log | cut -c1-600
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-400
timeout 300 python tools/update_sweep.py > gpurun_out/update_sweep.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/update_sweep.log
RED="--batch 16 --diffusion-steps 3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-roofline"
timeout 600 python bench.py $RED > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py $RED > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
CHUNK=64 ITERS=2 timeout 300 python tools/conv_one.py > gpurun_out/one.log 2>&1 && \
CHUNK=64 ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc4_kernel -s 1 -c 1 -o gpurun_out/prof_conv4_128_r1f -f python tools/conv_one.py > gpurun_out/ncu_one.log 2>&1; echo "ncu conv rc=$?"
B=64 timeout 300 python tools/update_one.py > gpurun_out/upd_one.log 2>&1 && \
B=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:superpose_update_kernel -s 12 -c 1 -o gpurun_out/prof_update_r1f -f python tools/update_one.py > gpurun_out/ncu_upd.log 2>&1; echo "ncu upd rc=$?"
cat gpurun_out/one.log gpurun_out/upd_one.log

#!/bin/bash
# Round evidence run (1 GPU): full GPU test suite, bench (own arm + reference arm), c5 update sweep, ncu launch lists
# (reference architecture and the attention-variant extension), ncu --set full of the two roofline kernels.
# Every ncu pass runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl gpurun_out/parity_growth.json
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
grep -h "TRAJECTORY\|SAME-SEED" gpurun_out/pytest_gpu.log | cut -c1-700
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_full.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-400
timeout 600 python bench.py --steps 2 --warmup 3 --arch attn > gpurun_out/bench_attn.log 2>&1; echo "bench attn rc=$?"; tail -1 gpurun_out/bench_attn.log | cut -c1-300
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --batch 8 > gpurun_out/bench_b8.log 2>&1; echo "bench b8 rc=$?"; tail -1 gpurun_out/bench_b8.log | cut -c1-250
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --res 128 --diffusion-steps 100 --batch 16 > gpurun_out/bench_c2.log 2>&1; echo "bench c2 rc=$?"; tail -1 gpurun_out/bench_c2.log | cut -c1-250
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --res 512 --diffusion-steps 1000 --batch 4 > gpurun_out/bench_c4.log 2>&1; echo "bench c4 shard rc=$?"; tail -1 gpurun_out/bench_c4.log | cut -c1-250
timeout 300 python tools/update_sweep.py > gpurun_out/update_sweep.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/update_sweep.log
timeout 300 python tools/conv_layers.py > gpurun_out/conv_layers.log 2>&1; cat gpurun_out/conv_layers.log
timeout 300 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; cat gpurun_out/attn_bench.log
RED="--diffusion-steps 3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-roofline"
timeout 600 python bench.py $RED > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 125 --csv --log-file gpurun_out/launches_ref.csv python bench.py $RED > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches(ref) rc=$?"
timeout 600 python bench.py $RED --arch attn > gpurun_out/plain_attn.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1130 -c 360 --csv --log-file gpurun_out/launches_attn.csv python bench.py $RED --arch attn > gpurun_out/ncu_launches_attn.log 2>&1; echo "ncu launches(attn) rc=$?"
CHUNK=64 ITERS=2 timeout 300 python tools/conv_one.py > gpurun_out/one.log 2>&1 && \
CHUNK=64 ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc4_kernel -s 3 -c 1 -o gpurun_out/prof_conv_128_r2 -f python tools/conv_one.py > gpurun_out/ncu_one.log 2>&1; echo "ncu conv rc=$?"
CIN=64 COUT=64 CHUNK=64 ITERS=2 timeout 300 python tools/conv_one.py > gpurun_out/one64.log 2>&1 && \
CIN=64 COUT=64 CHUNK=64 ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc4_kernel -s 3 -c 1 -o gpurun_out/prof_conv_64_r2 -f python tools/conv_one.py > gpurun_out/ncu_one64.log 2>&1; echo "ncu conv64 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel -s 5 -c 1 -o gpurun_out/prof_attn_r2 -f python tools/attn_bench.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
B=64 timeout 300 python tools/update_one.py > gpurun_out/upd_one.log 2>&1 && \
B=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:superpose_update_kernel -s 12 -c 1 -o gpurun_out/prof_update_r2 -f python tools/update_one.py > gpurun_out/ncu_upd.log 2>&1; echo "ncu upd rc=$?"
# DRAM traffic of the update step INCLUDING write-back: 20 launches deep inside the rotating sequence (buffer sets >> L2, so
# every launch evicts as many dirty lines as it produces), caches NOT flushed by the profiler, two counters = one pass
B=64 ITERS=200 timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:superpose_update_kernel -s 100 -c 20 --csv --log-file gpurun_out/upd_traffic_rotating.csv python tools/update_one.py > gpurun_out/ncu_upd2.log 2>&1; echo "ncu upd traffic rc=$?"
cat gpurun_out/one.log gpurun_out/upd_one.log

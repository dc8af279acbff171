#!/bin/bash
# experiment of the moment: conv layer A/B (fp32-math vs fp16-math transform), update sweep, half-math error, short bench
mkdir -p gpurun_out
HALF=$PWD/tools/_lib_half.so
python tools/conv_layers.py > gpurun_out/layers_f32_1.log 2>&1; cat gpurun_out/layers_f32_1.log
SDD_LIB=$HALF python tools/conv_layers.py > gpurun_out/layers_half_1.log 2>&1; cat gpurun_out/layers_half_1.log
python tools/conv_layers.py > gpurun_out/layers_f32_2.log 2>&1; tail -5 gpurun_out/layers_f32_2.log
SDD_LIB=$HALF python tools/conv_layers.py > gpurun_out/layers_half_2.log 2>&1; tail -5 gpurun_out/layers_half_2.log
CHUNK=8 python tools/conv_layers.py > gpurun_out/layers_f32_c8.log 2>&1; tail -5 gpurun_out/layers_f32_c8.log
BS=1,16,64,128,256 python tools/update_sweep.py > gpurun_out/update_sweep_r2a.log 2>&1; cat gpurun_out/update_sweep_r2a.log
SDD_LIB=$HALF timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trajectory.py -m gpu -q -p no:cacheprovider -k "unet_forward or c2_full or conv3x3_fused" > gpurun_out/pytest_half.log 2>&1; tail -3 gpurun_out/pytest_half.log; grep "TRAJECTORY" gpurun_out/pytest_half.log | cut -c1-900
grep -h '"unet_forward"' gpurun_out/parity_report.jsonl | tail -8
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2b.log 2>&1; head -c 330 gpurun_out/bench_r2b.log; echo
SDD_LIB=$HALF timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2b_half.log 2>&1; head -c 330 gpurun_out/bench_r2b_half.log; echo

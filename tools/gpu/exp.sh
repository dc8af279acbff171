#!/bin/bash
# experiment of the moment: N2 tests, fp16 vs bf16 A/B on the conv layers, small-batch bench, extension bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet_attn.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "attn or attention" > gpurun_out/pytest_attn.log 2>&1; tail -15 gpurun_out/pytest_attn.log | cut -c1-300
grep -h "unet_attn" gpurun_out/parity_report.jsonl | tail -8
BF=$PWD/tools/_lib_bf16.so
python tools/conv_layers.py > gpurun_out/layers_f16_a.log 2>&1; cat gpurun_out/layers_f16_a.log
SDD_LIB=$BF python tools/conv_layers.py > gpurun_out/layers_bf16_a.log 2>&1; cat gpurun_out/layers_bf16_a.log
python tools/conv_layers.py > gpurun_out/layers_f16_b.log 2>&1; tail -5 gpurun_out/layers_f16_b.log
SDD_LIB=$BF python tools/conv_layers.py > gpurun_out/layers_bf16_b.log 2>&1; tail -5 gpurun_out/layers_bf16_b.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --batch 8 > gpurun_out/bench_b8.log 2>&1; head -c 330 gpurun_out/bench_b8.log; echo
timeout 600 python bench.py --steps 2 --warmup 3 --arch attn > gpurun_out/bench_attn.log 2>&1; tail -c 2500 gpurun_out/bench_attn.log; echo
timeout 300 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; cat gpurun_out/attn_bench.log

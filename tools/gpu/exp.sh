#!/bin/bash
# experiment of the moment: attention kernel with two CTAs per SM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet_attn.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "attn or attention" > gpurun_out/pytest_attn.log 2>&1; tail -3 gpurun_out/pytest_attn.log | cut -c1-300
timeout 300 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; cat gpurun_out/attn_bench.log
timeout 600 python bench.py --steps 2 --warmup 3 --arch attn > gpurun_out/bench_attn.log 2>&1; head -c 300 gpurun_out/bench_attn.log; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel -s 3 -c 1 -o gpurun_out/prof_attn_r2 -f python tools/attn_bench.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"

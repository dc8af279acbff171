#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x -k "${1:-}" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 1 --warmup 1 --diffusion-steps 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_short.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_short.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('samples/s(T=10)',d['value'], 'frac', d['whole_path_tensor_frac_of_sustained'], 'conv', d['roofline']['frac'], 'upd', d['roofline_update']['frac'], 'launches', d['gpu_launches'])"
grep PARITY gpurun_out/pytest_gpu.log | grep -E "unet_forward|ddpm_sample|superposed" | tail -12

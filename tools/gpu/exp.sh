#!/bin/bash
# experiment of the moment: validate TMA-staged conv_out1 + resample / projection fixes; timings
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_gpu_unet_attn.py tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_exp.log 2>&1; tail -6 gpurun_out/pytest_exp.log | cut -c1-300
grep -h "unet_attn\|\"unet_forward\"" gpurun_out/parity_report.jsonl | tail -14
RED="--diffusion-steps 3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-roofline"
timeout 600 python bench.py $RED > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 395 -c 125 --csv --log-file gpurun_out/launches_ref.csv python bench.py $RED > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches(ref) rc=$?"
timeout 600 python bench.py $RED --arch attn > gpurun_out/plain_attn.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1130 -c 360 --csv --log-file gpurun_out/launches_attn.csv python bench.py $RED --arch attn > gpurun_out/ncu_launches_attn.log 2>&1; echo "ncu launches(attn) rc=$?"
python tools/launch_summary.py gpurun_out/launches_ref.csv gpurun_out/launches_attn.csv
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_exp.log 2>&1; head -c 330 gpurun_out/bench_exp.log; echo
timeout 600 python bench.py --steps 2 --warmup 3 --arch attn --no-e2e > gpurun_out/bench_attn.log 2>&1; head -c 330 gpurun_out/bench_attn.log; echo

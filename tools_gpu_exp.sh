#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --batch 16 --res 128 --diffusion-steps 100 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1; echo "c2 rc=$?"; tail -1 gpurun_out/bench_c2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline_update']['frac'], d['config']['workload'])"
timeout 600 python bench.py --batch 4 --res 512 --diffusion-steps 1000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.log 2>&1; echo "c4 rc=$?"; tail -1 gpurun_out/bench_c4.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline_update']['frac'], d['config']['workload'])"
timeout 300 python __graft_entry__.py smoke

#!/bin/bash
# does the fp16-math transform build hold every bound of the test suite?  (+ same-box bench A/B)
mkdir -p gpurun_out
H=$PWD/tools/_lib_half.so
rm -f gpurun_out/parity_report.jsonl gpurun_out/parity_growth.json
SDD_LIB=$H timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_half.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_half.log | cut -c1-300
cp gpurun_out/parity_report.jsonl gpurun_out/parity_report_half.jsonl; cp gpurun_out/parity_growth.json gpurun_out/parity_growth_half.json
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_f32m.log 2>&1; head -c 250 gpurun_out/bench_f32m.log; echo
SDD_LIB=$H timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_halfm.log 2>&1; head -c 250 gpurun_out/bench_halfm.log; echo
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_f32m_2.log 2>&1; head -c 250 gpurun_out/bench_f32m_2.log; echo
SDD_LIB=$H timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_halfm_2.log 2>&1; head -c 250 gpurun_out/bench_halfm_2.log; echo

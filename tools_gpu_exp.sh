#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -x -k "unet_forward or k3 or k5 or ddpm_sample or baseline_res or shard" > gpurun_out/pytest_o1.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_o1.log
RED="--batch 16 --diffusion-steps 3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
for CPS in 2 1; do
SDD_CIN_CPS=$CPS timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'conv_out1|conv_in_mma|conv_out2' -s 12 -c 24 --csv --log-file gpurun_out/o1_$CPS.csv python bench.py $RED > gpurun_out/ncu_o1.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv,collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/o1_$CPS.csv') if not l.startswith('=='))]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
agg=collections.defaultdict(list)
for r in rows[1:]:
    if len(r)>vi:
        v=float(r[vi].replace(',','')); u=r[ui]
        v = v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
        agg[r[ki][:40]].append(v)
for k,v in agg.items(): print('CPS=$CPS',k,len(v),'avg us',sum(v)/len(v))
PY
done

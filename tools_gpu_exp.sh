#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; grep -v "timed out" gpurun_out/pytest_gpu.log | tail -3
grep PARITY gpurun_out/pytest_gpu.log | grep -E "unet_forward|ddpm_sample|superposed" | tail -12
if [ $rc -ne 0 ]; then exit 1; fi
RED="--batch 16 --diffusion-steps 3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 python bench.py $RED > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py $RED > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; c={n:i for i,n in enumerate(h)}
agg=collections.defaultdict(list)
for r in rows[hi+1:]:
    if len(r)==len(h) and r[c['Metric Name']]=='gpu__time_duration.sum':
        v=float(r[c['Metric Value']].replace(',','')); u=r[c['Metric Unit']]
        v = v/1000 if u in ('ns','nsecond') else v
        agg[r[c['Kernel Name']].split('(')[0]].append(v)
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print(f"{k[:60]:60s} n={len(v):3d} total={sum(v):9.1f} us avg={sum(v)/len(v):8.1f} share={100*sum(v)/tot:5.1f}%")
PY

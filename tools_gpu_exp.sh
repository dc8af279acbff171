#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 120 -x -k "update or philox or k3 or shard or superpos" > gpurun_out/pytest_upd.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_upd.log
for CPS in 2 3; do
echo "ring kernel, CTAs/SM = $CPS"
SDD_UPD_CPS=$CPS MODE=rotating BS=1,16,64,128,256,512 OUT=gpurun_out/upd_ring_$CPS.json timeout 200 python tools/update_sweep.py 2>&1 | grep -v "timed out"
done
SDD_UPD_CPS=2 MODE=flush BS=64,256 OUT=gpurun_out/upd_ring_flush.json timeout 200 python tools/update_sweep.py 2>&1 | grep -v "timed out"

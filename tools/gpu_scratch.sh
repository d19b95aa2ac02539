#!/bin/bash
# four-step chunk size experiment (SP_SCRATCH_MB): C5 / C3-z1 / C4 n=16384
OUT=gpurun_out; mkdir -p $OUT
for mb in 0 160 320 640 1280; do
  echo "== SP_SCRATCH_MB=$mb"
  SP_SCRATCH_MB=$mb timeout 300 python tools/sweep.py C5,C3-z1,C3-z4 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['case'],d['fmt'],d['n'],'ms',round(d['ms_per_render'],3),'GS/s',round(d['msamples_s']/1e3,1),'frac',round(d['frac_of_measured_hbm'],3),'launches',d['launches'])"
done | tee $OUT/scratch_$1.txt

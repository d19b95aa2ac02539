#!/bin/bash
# performance experiments: which part of the r64 kernel costs what (SP_DEBUG_SKIP: 1 stores, 2 atomics, 4 LUT)
OUT=gpurun_out; mkdir -p $OUT
for m in 0 1 4 5; do
  SP_DEBUG_SKIP=$m timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('skip=$m kernel_ms',d['roofline']['kernel_ms'],'frac',round(d['roofline']['frac'],4))"
done | tee $OUT/dbg_$1.txt

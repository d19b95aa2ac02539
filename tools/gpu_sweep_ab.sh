#!/bin/bash
# A/B of experiment builds over a list of tools/sweep.py cases: usage (under gpurun) bash tools/gpu_sweep_ab.sh TAG CASES lib_a lib_b ...
TAG=$1; CASES=$2; shift 2; OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do for name in "$@"; do
  SP_LIB=$(realpath spectroplot-js_b200/$name)/libspectro_b200.so timeout 300 python tools/sweep.py $CASES 5 2>>$OUT/sweepab_$TAG.err | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('$name', d['case'], 'ms %.4f GS/s %.1f ok=%s' % (d['ms_per_render'], d['msamples_s'] / 1e3, d['hist_ok']))" | tee -a $OUT/sweepab_$TAG.txt
done; done

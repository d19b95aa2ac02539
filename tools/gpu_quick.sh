#!/bin/bash
# quick perf iteration: parity tests + bench (no cpu leg) [+ optional ncu full] [+ microbench]
# usage (under gpurun): bash tools/gpu_quick.sh TAG [ncu] [micro]
TAG=${1:-q}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?" >> $OUT/bench_$TAG.err
if [[ " $* " == *" ncu "* ]]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ -s 3 -c 1 -f -o $OUT/prof_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full_$TAG.log 2>&1
fi
if [[ " $* " == *" micro "* ]]; then
timeout 300 tools/microbench/pipes > $OUT/pipes_$TAG.txt 2>&1
fi
tail -4 $OUT/pytest_$TAG.log; python -c "
import json;d=json.load(open('$OUT/bench_$TAG.json'));print('value',d['value'],'ms/step',d['ms_per_step'],'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['clocks'])"; tail -3 $OUT/bench_$TAG.err

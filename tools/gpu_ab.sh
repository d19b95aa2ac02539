#!/bin/bash
# A/B of experiment builds of the C library on one GPU box: for every lib_*/ under spectroplot-js_b200 (csrc/Makefile XFLAGS)
# the device-resident C2 leg of bench.py (kernel ms, step ms); libs listed in $TEST_LIBS also run the GPU parity suite.
# usage (under gpurun): TEST_LIBS="lib_d" bash tools/gpu_ab.sh TAG
TAG=${1:-ab}; OUT=gpurun_out; mkdir -p $OUT
: > $OUT/ab_$TAG.jsonl
for rep in 1 2; do
for d in spectroplot-js_b200/lib_*/; do
  lib=$(realpath $d)/libspectro_b200.so
  [ -f $lib ] || continue
  SP_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 3 --kernel-only >> $OUT/ab_$TAG.jsonl 2>> $OUT/ab_$TAG.err
done
done
for name in $TEST_LIBS; do
  lib=$(realpath spectroplot-js_b200/$name)/libspectro_b200.so
  SP_LIB=$lib timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_${name}_$TAG.log 2>&1; echo "rc=$?" >> $OUT/pytest_${name}_$TAG.log
  tail -3 $OUT/pytest_${name}_$TAG.log
done
python - <<PY
import json
for l in open("$OUT/ab_$TAG.jsonl"):
    d = json.loads(l); print(d["lib"].split("/")[-2], "kernel_ms %.4f step %.4f frac %.4f ok=%s" % (d["kernel_ms"], d["ms_per_step"], d["frac"], d["hist_ok"]))
PY

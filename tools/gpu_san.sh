#!/bin/bash
# compute-sanitizer over the kernels that are new or changed in round 2.  usage (under gpurun): bash tools/gpu_san.sh TAG
TAG=${1:-r2}; OUT=gpurun_out; mkdir -p $OUT
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/san_multi.py 0 1 9 12 13 14 15 16 17 18 19 20 21 22 > $OUT/san_${tool}_$TAG.log 2>&1; echo "rc=$?" >> $OUT/san_${tool}_$TAG.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|rc=" $OUT/san_${tool}_$TAG.log | tail -10
done

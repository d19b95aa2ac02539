#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list, ncu full capture of the render kernel.
# usage (under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt
nproc > $OUT/nproc_$TAG.txt
timeout -s KILL 300 python -m pytest tests -m gpu -q --durations=5 > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
timeout -s KILL 200 python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?" >> $OUT/bench_$TAG.err
timeout -s KILL 200 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launch_$TAG.log 2>&1
timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k regex:render_ -s 3 -c 1 -f -o $OUT/prof_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full_$TAG.log 2>&1
tail -3 $OUT/pytest_gpu_$TAG.log; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err; ls -la $OUT

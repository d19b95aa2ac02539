#!/bin/bash
# Evidence run: (1) DRAM / L2 counters of the render kernel on overlapping hops (C2 at zoom 1, 2, 4, 8: frames share
# n - H samples; the re-read must come from L2, not HBM), (2) compute-sanitizer over every kernel family.
# usage (under gpurun): bash tools/gpu_evidence.sh TAG
TAG=${1:-e1}; OUT=gpurun_out; mkdir -p $OUT
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_read_lookup_hit.sum,lts__t_sectors_op_read_lookup_miss.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:render_r64 --csv --log-file $OUT/overlap_$TAG.csv \
    python tools/sweep.py C2-hann,C2-z2,C2-z4,C2-z8 1 > $OUT/overlap_sweep_$TAG.log 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/san_multi.py > $OUT/san_${tool}_$TAG.log 2>&1; echo "rc=$?" >> $OUT/san_${tool}_$TAG.log
done
for f in $OUT/san_*_$TAG.log; do tail -n 3 $f; done; wc -l $OUT/overlap_$TAG.csv

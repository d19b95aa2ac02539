#!/bin/bash
# hop-overlap counters of render_w_kernel with (SP_DEBUG_SKIP=0) and without (16) span staging, plus a re-check of the
# big kernel's DRAM traffic.  usage (under gpurun): bash tools/gpu_overlap.sh TAG
TAG=${1:-r2}; OUT=gpurun_out; mkdir -p $OUT
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sector_op_read_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed.sum
for dbg in 0 16; do
  SP_DEBUG_SKIP=$dbg timeout -s KILL 400 ncu --metrics $M --clock-control none -k regex:render_w_ -c 40 --csv --log-file $OUT/overlap_w_dbg${dbg}_$TAG.csv \
      python tools/sweep.py X:CS16:128:1:24,X:CS16:128:2:24,X:CS16:128:4:24,X:CS16:128:8:24,X:CS16:256:4:24,X:CS16:512:4:24 1 > $OUT/overlap_w_dbg${dbg}_$TAG.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:render_big -s 3 -c 1 --csv --log-file $OUT/ncu_big_traffic_$TAG.csv python tools/sweep.py C5 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:render_big -s 3 -c 1 --csv --log-file $OUT/ncu_big_traffic_c3_$TAG.csv python tools/sweep.py C3-z1 1 > /dev/null 2>&1
python tools/sweep.py C5,C3-z1,X:CS16:128:1:26,X:CS16:256:1:26,X:CS16:512:1:26,X:CS16:1024:1:26,X:CU8:512:1:26 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('%s ms %.3f GS/s %.1f frac %.3f' % (d['case'], d['ms_per_render'], d['msamples_s']/1e3, d['frac_of_measured_hbm']))"
tail -4 $OUT/ncu_big_traffic_$TAG.csv | cut -d, -f5,13-; tail -3 $OUT/ncu_big_traffic_c3_$TAG.csv | cut -d, -f5,13-

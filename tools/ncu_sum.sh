#!/bin/bash
# usage: tools/ncu_sum.sh gpurun_out/prof_TAG.ncu-rep  -> prints the summary (raw metrics + opcode mix)
REP=$1; B=/tmp/ncu_$(basename $REP .ncu-rep)
ncu -i $REP --page raw --csv > $B.raw.csv 2>/dev/null
ncu -i $REP --page source --csv > $B.src.csv 2>/dev/null
python $(dirname $0)/ncu_summary.py $B.raw.csv $B.src.csv

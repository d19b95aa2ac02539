#!/bin/bash
# A/B of render_w_kernel builds (spectroplot-js_b200/lib_w*): cs16 / cu8 at N = 64 .. 1024, hop N and zoom 4.
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/wtune.txt
for rep in 1 2; do
for d in spectroplot-js_b200/lib_w*/; do
  lib=$(realpath $d)/libspectro_b200.so
  SP_LIB=$lib timeout 100 python tools/sweep.py X:CS16:64:1:26,X:CS16:128:1:26,X:CS16:256:1:26,X:CS16:512:1:26,X:CS16:1024:1:26,X:CU8:512:1:26,X:CS16:128:4:24,X:CS16:512:4:24 5 2>/dev/null | python -c "
import json,sys
print('$(basename $d)', ' '.join('%s:%.0f' % (json.loads(l)['case'][2:], json.loads(l)['msamples_s']/1e3) for l in sys.stdin))" >> $OUT/wtune.txt
done; done
cat $OUT/wtune.txt

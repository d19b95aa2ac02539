#!/bin/bash
# timing of one experiment build ($1 = lib dir name under spectroplot-js_b200) over a list of SP_DEBUG_SKIP values ($2...)
LIB=$(realpath spectroplot-js_b200/$1)/libspectro_b200.so; shift
for rep in 1 2; do for d in "$@"; do
  echo -n "dbg=$d "; SP_DEBUG_SKIP=$d SP_LIB=$LIB timeout 200 python bench.py --steps 20 --warmup 3 --kernel-only 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kernel_ms %.4f step %.4f' % (d['kernel_ms'], d['ms_per_step']))"
done; done

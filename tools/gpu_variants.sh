#!/bin/bash
# compare kernel variants: SP_NO_BIG (render_kernel) and SP_BIG_VARIANT=0..3 (needs -DSP_BIG_VARIANTS build)
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_var.log 2>&1; tail -2 $OUT/pytest_var.log
for v in nobig 0 1 2 3; do
  if [ $v = nobig ]; then export SP_NO_BIG=1; else unset SP_NO_BIG; export SP_BIG_VARIANT=$v; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench_var_$v.json 2> $OUT/bench_var_$v.err
  python -c "
import json;d=json.load(open('$OUT/bench_var_$v.json'));print('variant $v: kernel_ms %.4f ms/step %.4f frac %.4f'%(d['roofline']['kernel_ms'],d['ms_per_step'],d['roofline']['frac']))" || tail -3 $OUT/bench_var_$v.err
done

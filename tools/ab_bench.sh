#!/bin/bash
# A/B bench of library variants under build_ab/ (runs on the GPU box); TICK_FLAGS = extra --tick-flags
mkdir -p gpurun_out
run() { name=$1; shift
  out=$(env "$@" python bench.py --steps 400 --warmup 50 --no-cpu-baseline --e2e-steps 16 --tick-flags ${TICK_FLAGS:-0} 2>&1 | tail -1)
  echo "$name $(echo "$out" | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('ms/step %.4f frac %.3f value %.4g e2e %.4g' % (d['ms_per_step'], d['roofline']['frac'], d['value'], d['e2e']['value']))
except Exception as e: print('FAILED', e)")" | tee -a gpurun_out/ab.log
}
run default X=1
for f in build_ab/*.so; do run "$(basename $f)" ASTRO_B200_LIB=$PWD/$f; done

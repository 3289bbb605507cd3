#!/bin/bash
# A/B bench of library variants under build_ab/ (runs on the GPU box). usage: tools/ab_bench.sh [extra bench args]
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  out=$(env "$@" python bench.py --steps 400 --warmup 50 --no-cpu-baseline --e2e-steps 20 2>&1 | tail -1)
  echo "$name $(echo "$out" | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('ms/step %.4f frac %.3f value %.4g e2e %.4g' % (d['ms_per_step'], d['roofline']['frac'], d['value'], d['e2e']['value']))
except Exception as e: print('FAILED', e)")" | tee -a gpurun_out/ab.log
}
run default X=1
run gran32 ASTRO_L2_FETCH_GRANULARITY=32
run gran128 ASTRO_L2_FETCH_GRANULARITY=128
for f in build_ab/*.so; do run "$(basename $f)" ASTRO_B200_LIB=$PWD/$f; done
for f in build_ab/*.so; do run "$(basename $f)+gran32" ASTRO_B200_LIB=$PWD/$f ASTRO_L2_FETCH_GRANULARITY=32; done

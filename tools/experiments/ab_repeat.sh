#!/bin/bash
# noise check: every library (default + build_ab/*.so) REPS times, interleaved (GPU box)
run() { name=$1; shift
  out=$(env "$@" python bench.py --steps 400 --warmup 50 --no-cpu-baseline --e2e-steps 4 2>&1 | tail -1)
  echo "$name $(echo "$out" | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('ms/step %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))
except Exception as e: print('FAILED', e)")" | tee -a gpurun_out/ab.log
}
for r in $(seq ${REPS:-3}); do
  run default X=1
  for f in build_ab/*.so; do run "$(basename $f)" ASTRO_B200_LIB=$PWD/$f; done
done

#!/bin/bash
# A/B of libraries (default + build_ab/*.so): fused value and per-tick-launch value of bench.py (GPU box)
run() { name=$1; shift
  out=$(env "$@" python bench.py --steps 640 --warmup 64 --no-cpu-baseline --no-rollout --no-fresh --strong-total 0 --e2e-steps 4 2>&1 | tail -1)
  echo "$name $(echo "$out" | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('fused us/tick %.2f  per-tick-launch us/tick %.2f' % (1e3*d['ms_per_step'], 1e3*d['per_tick_launch']['ms_per_step']))
except Exception as e: print('FAILED', e)")" | tee -a gpurun_out/ab.log
}
for r in $(seq ${REPS:-1}); do
  run default X=1
  for f in build_ab/*.so; do run "$(basename $f)" ASTRO_B200_LIB=$PWD/$f; done
done

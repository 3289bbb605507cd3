#!/bin/bash
# A/B at the strong-scaled shape (BASELINE configs[3]: 131,072 games per GPU) and at 1,048,576: default library + build_ab/*.so
run() { name=$1; games=$2; shift; shift
  out=$(env "$@" python bench.py --games-per-gpu $games --steps 20 --warmup 5 --no-cpu-baseline --no-rollout --no-fresh --strong-total 0 --e2e-steps 4 2>&1 | tail -1)
  echo "$name games=$games $(echo "$out" | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('us/tick %.2f  value %.4g  per-tick-launch %.2f' % (1e3*d['ms_per_step'], d['value'], 1e3*d['per_tick_launch']['ms_per_step']))
except Exception as e: print('FAILED', e)")" | tee -a gpurun_out/ab_small.log
}
for games in ${GAMES:-131072 262144}; do
  run default $games X=1
  for f in build_ab/*.so; do if [ -f $f ]; then run "$(basename $f)" $games ASTRO_B200_LIB=$PWD/$f; fi; done
done

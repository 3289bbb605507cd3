#!/bin/bash
# A/B of environment knobs and libraries (GPU box): bench.py at the driver's shape (--steps 20: ONE launch of 20 ticks) and at 64 ticks
# per launch.  usage: tools/ab_env.sh "<VAR=VALUE[,VAR=VALUE...]> ..."   (build_ab/*.so are run too, with the default environment)
run() { name=$1; steps=$2; shift; shift
  out=$(env "$@" python bench.py --steps $steps --warmup 5 --no-cpu-baseline --no-rollout --no-fresh --strong-total 0 --e2e-steps 4 2>&1 | tail -1)
  echo "$name steps=$steps $(echo "$out" | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('us/tick %.2f  value %.4g  per-tick-launch %.2f  episodes %d' % (1e3*d['ms_per_step'], d['value'], 1e3*d['per_tick_launch']['ms_per_step'], d['episode_stats']['episodes']))
except Exception as e: print('FAILED', e)")" | tee -a gpurun_out/ab_env.log
}
for r in $(seq ${REPS:-1}); do
  for steps in ${STEPS:-20 64}; do
    for f in build_ab/*.so; do if [ -f $f ]; then run "$(basename $f)" $steps ASTRO_B200_LIB=$PWD/$f; fi; done
    for cfg in $1; do
      run "$cfg" $steps $(echo $cfg | tr ',' ' ')
    done
  done
done

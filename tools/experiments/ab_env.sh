#!/bin/bash
# A/B of environment knobs with the default library (runs on the GPU box)
run() { name=$1; shift
  out=$(env "$@" python bench.py --steps 400 --warmup 50 --no-cpu-baseline --e2e-steps 16 2>&1 | tail -1)
  echo "$name $(echo "$out" | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('ms/step %.4f frac %.3f value %.4g e2e %.4g' % (d['ms_per_step'], d['roofline']['frac'], d['value'], d['e2e']['value']))
except Exception as e: print('FAILED', e)")" | tee -a gpurun_out/ab.log
}
for g in 32 64 128; do run gran$g ASTRO_L2_FETCH_GRANULARITY=$g; done

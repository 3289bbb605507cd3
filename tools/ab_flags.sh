#!/bin/bash
# A/B of kernel variants inside one library via --tick-flags (runs on the GPU box)
mkdir -p gpurun_out
for f in 0 8 16; do
  for steps in 500; do
    out=$(python bench.py --steps $steps --warmup 50 --no-cpu-baseline --e2e-steps 20 --tick-flags $f 2>&1 | tail -1)
    echo "flags=$f steps=$steps $(echo "$out" | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('ms/step %.4f frac %.3f value %.4g e2e %.4g clocks %s' % (d['ms_per_step'], d['roofline']['frac'], d['value'], d['e2e']['value'], d['clocks']))
except Exception as e: print('FAILED', e)")" | tee -a gpurun_out/ab_flags.log
  done
done

#!/bin/bash
# Runs on the GPU box (under gpurun): GPU tests, a bench line, then an ncu capture of the tick kernel.
# usage: tools/gpu_check.sh <tag> [skip-tests]
TAG=${1:-run}
mkdir -p gpurun_out
if [ "$2" != "skip-tests" ]; then
  timeout 1300 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
fi
timeout 300 python bench.py --steps 500 --warmup 50 > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_$TAG.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.4g  ms/step %.4f  frac %.3f  B/step %.1f  e2e %.4g  cpu %.3g' % (d['value'], d['ms_per_step'], r['frac'], r['bytes_per_env_step'], d['e2e']['value'], (d['cpu_baseline'] or {}).get('value', 0)))"
# the profiled launches sit in the stationary population: 1,200 pre-roll ticks (one launch each), 1 warm-up launch, then
# the timed launches of astro_tick_many (index 1201: 64 ticks of every tile) and the one-launch-per-tick loop (index 1222)
CMD="python bench.py --steps 64 --warmup 5 --preroll 1200 --no-cpu-baseline --e2e-steps 8"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tick_f32_kernel -s 1222 -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tick_f32_kernel -s 1201 -c 1 -f -o gpurun_out/prof_${TAG}_fused $CMD > gpurun_out/ncu_${TAG}_fused.log 2>&1; echo "ncu fused rc=$?"

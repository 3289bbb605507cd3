#!/bin/bash
# dynamic instruction counts of library variants: the timed launches of a short bench run under ncu (GPU box);
# raw CSVs land in gpurun_out/abncu_<variant>.csv (read them with tools/ab_ncu_read.py)
CMD="python bench.py --steps 64 --warmup 64 --preroll 300 --no-cpu-baseline --no-rollout --no-fresh --strong-total 0 --e2e-steps 4"
for f in default build_ab/*.so; do
  name=$(basename $f .so)
  if [ "$f" = default ]; then E="X=1"; else E="ASTRO_B200_LIB=$PWD/$f"; fi
  env $E ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none \
      -k regex:tick_f32_kernel -s 301 -c 4 --csv --log-file gpurun_out/abncu_$name.csv $CMD > /dev/null 2>&1
  echo "$name rc=$?"
done

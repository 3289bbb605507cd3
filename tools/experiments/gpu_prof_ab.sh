#!/bin/bash
# ncu capture of the tick kernel for the default library and every build_ab variant (GPU box)
CMD="python bench.py --steps 40 --warmup 5 --preroll 1200 --no-cpu-baseline --e2e-steps 8"
prof() { name=$1; shift
  env "$@" ncu --set full --clock-control none --import-source on -k regex:tick_f32_kernel -s 1230 -c 1 -f -o gpurun_out/prof_$name $CMD > gpurun_out/ncu_$name.log 2>&1; echo "$name ncu rc=$?"; }
prof ${TAG:-cur}_default X=1
for f in build_ab/*.so; do prof ${TAG:-cur}_$(basename $f .so) ASTRO_B200_LIB=$PWD/$f; done

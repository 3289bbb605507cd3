#!/bin/bash
# Secondary-kernel measurements on the GPU box: observe, config #5 rollouts, bot loops (see profiles/r1_secondary_kernels.md)
mkdir -p gpurun_out
{
echo "== bench_observe"; timeout 200 python tools/bench_observe.py
echo "== bench_rollout"; timeout 300 python tools/bench_rollout.py
echo "== bench_rollout --shared"; timeout 300 python tools/bench_rollout.py --shared
echo "== bench_rollout --fused"; timeout 300 python tools/bench_rollout.py --fused
echo "== bench_bots"; timeout 300 python tools/bench_bots.py
} > gpurun_out/secondary_${1:-run}.log 2>&1
tail -40 gpurun_out/secondary_${1:-run}.log

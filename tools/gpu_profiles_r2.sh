#!/bin/bash
# Round-2 evidence run (GPU box, under gpurun): a plain bench run, the ncu launch list of the same command, and ncu --set full
# captures of the tick kernel at the launch shapes bench.py times (20 ticks per launch = the driver's --steps 20; 64; 1) and of
# the policy kernel.  Everything lands in gpurun_out/; tools/save_profile.py turns the reports into profiles/r2_*.
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-rollout --no-fresh --strong-total 0"
$B > gpurun_out/r2b_bench_plain.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2b_launches.csv $B > gpurun_out/r2b_launches.log 2>&1; echo "launch list rc=$?"
NC="ncu --set full --clock-control none --import-source on --kernel-name-base mangled -f"
$NC -k regex:tick_f32_kernelILi2ELb1ELb1E -s 1 -c 1 -o gpurun_out/r2b_tick_fused20 $B > /dev/null 2>&1; echo "fused20 rc=$?"
$NC -k regex:tick_f32_kernelILi2ELb1ELb1E -s 1 -c 1 -o gpurun_out/r2b_tick_fused64 python bench.py --steps 64 --warmup 64 --no-cpu-baseline --no-rollout --no-fresh --strong-total 0 > /dev/null 2>&1; echo "fused64 rc=$?"
$NC -k regex:tick_f32_kernelILi2ELb1ELb0E -s 640 -c 1 -o gpurun_out/r2b_tick_single $B > /dev/null 2>&1; echo "single rc=$?"
$NC -k regex:policy_mma -s 3 -c 1 -o gpurun_out/r2b_policy_mma python tools/prof_policy.py > /dev/null 2>&1; echo "policy rc=$?"
ls -la gpurun_out/r2b_*.ncu-rep
$NC -k regex:value_forward -s 3 -c 1 -o gpurun_out/r2b_value_forward python tools/prof_value.py > /dev/null 2>&1; echo "value_forward rc=$?"

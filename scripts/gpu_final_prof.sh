#!/bin/bash
# final profiles of the round: launch list of one eager step + ncu --set full of the MLP / weight-gradient kernels of one step.
# The .ncu-rep stays on the box (too large to bring back); its raw-metrics page is exported as CSV.
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 700 --csv --log-file gpurun_out/r02f_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-kernel-timing --no-extras > gpurun_out/r02f_launches.log 2>&1
echo launches rc=$?
timeout 600 ncu --set full --clock-control none -k regex:'tcs_|outer_accum_tc|sdf_only_tc4' -s 23 -c 23 -f -o /tmp/r02f_full \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-kernel-timing --no-extras > gpurun_out/r02f_full.log 2>&1
echo full rc=$?; ls -la /tmp/r02f_full.ncu-rep
ncu -i /tmp/r02f_full.ncu-rep --page raw --csv > gpurun_out/r02f_full_raw.csv 2>/dev/null; ls -la gpurun_out/r02f_full_raw.csv

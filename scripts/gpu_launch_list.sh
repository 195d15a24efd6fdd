#!/bin/bash
# ncu launch list (per-launch durations) of one eager demo_2 mapping step
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s ${2:-300} -c ${3:-330} --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-kernel-timing --no-extras > gpurun_out/${tag}_launches.log 2>&1
echo rc=$?; wc -l gpurun_out/${tag}_launches.csv

#!/bin/bash
# ncu capture (source-level stall sampling) of the SDF main-pass kernels of one eager step
mkdir -p gpurun_out
timeout 500 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section MemoryWorkloadAnalysis \
    --section LaunchStats --section Occupancy --section SchedulerStats --import-source on --clock-control none \
    -k regex:tcs -s 16 -c 16 -f -o gpurun_out/tcs_prof \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-kernel-timing --no-extras > gpurun_out/tcs_prof.log 2>&1
echo ncu rc=$?
ls -la gpurun_out/tcs_prof.ncu-rep

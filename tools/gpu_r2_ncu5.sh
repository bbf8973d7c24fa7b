#!/bin/bash
# the five launches of one bench pass: duration, instruction count, issue utilisation, warp-state breakdown
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --section SpeedOfLight --section SchedulerStats --section WarpStateStats --section InstructionStats --clock-control none -k regex:mc_forward -s 15 -c 5 -f -o gpurun_out/r2_five $CMD > gpurun_out/ncu5.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu5.log

#!/bin/bash
# ncu --set full (with source) of the five mc_forward launches of one timed bench pass, after a plain run of the same command
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux --no-extras"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mc_forward -s 15 -c 5 -f -o gpurun_out/r2_full5 $CMD > gpurun_out/ncu_full5.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_full5.log; tail -1 gpurun_out/plain.log | cut -c1-300

#!/bin/bash
# full ncu capture (with source) of the len = 8 launch of the bench pass
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mc_forward -s 19 -c 1 -f -o gpurun_out/r2_mc8 $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu.log

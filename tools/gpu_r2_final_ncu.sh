#!/bin/bash
# (1) ncu launch list of the default bench command, (2) ncu --set full of the fine-tune and top-k kernels; each after a plain run exited 0
mkdir -p gpurun_out
if [ "$1" = "list" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/plain_list.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "ncu exit $?"; tail -1 gpurun_out/plain_list.log | cut -c1-200; wc -l gpurun_out/r2_launches.csv
else
  CMD="python tools/other_kernels.py"
  $CMD > gpurun_out/plain_other.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"train_|topk_global" -s 8 -c 10 -f -o gpurun_out/r2_other $CMD > gpurun_out/ncu_other.log 2>&1
  echo "ncu exit $?"; tail -2 gpurun_out/ncu_other.log; cat gpurun_out/plain_other.log | tail -1
fi

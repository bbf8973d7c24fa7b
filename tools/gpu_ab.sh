#!/bin/bash
# A/B of library variants on the bench workload: tools/gpu_ab.sh default name1 name2 ...  (build/variants/libbfb_<name>.so)
mkdir -p gpurun_out
for v in "$@"; do
  if [ $v != default ]; then export BAYESFORMER_B200_LIB=$PWD/build/variants/libbfb_$v.so; else unset BAYESFORMER_B200_LIB; fi
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/ab_$v.log 2>&1; echo "bench $v exit $?"
  python - <<PY
import json
l=[x for x in open('gpurun_out/ab_$v.log') if x.startswith('{')][-1]
d=json.loads(l); print('$v', 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'launch ms', round(d['roofline']['avg_launch_ms'],3))
PY
done

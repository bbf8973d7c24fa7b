#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2b_fused_tests.log 2>&1; echo "fused tests exit $?"; tail -3 gpurun_out/r2b_fused_tests.log
timeout 600 python tools/fused_accuracy.py > gpurun_out/r2b_accuracy.log 2>&1; echo "accuracy exit $?"; tail -6 gpurun_out/r2b_accuracy.log
for v in default; do
  if [ $v != default ]; then export BAYESFORMER_B200_LIB=$PWD/build/variants/libbfb_$v.so; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/r2b_bench_$v.log 2>&1; echo "bench $v exit $?"
  python - <<PY
import json
l=[x for x in open('gpurun_out/r2b_bench_$v.log') if x.startswith('{')][-1]
d=json.loads(l); print('$v', 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'launch ms', d['roofline']['avg_launch_ms'])
PY
done

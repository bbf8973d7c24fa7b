#!/bin/bash
mkdir -p gpurun_out
run() { timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/ab_$1.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/ab_$1.log') if x.startswith('{')][-1]
d=json.loads(l); print('$1', 'value', round(d['value']), 'ms', round(d['ms_per_step'],2))
PY
}
run default
BFB200_FUSED_FORCE_LONG=1 run long

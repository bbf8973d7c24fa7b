#!/bin/bash
mkdir -p gpurun_out
run() { timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/ab_$1.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/ab_$1.log') if x.startswith('{')][-1]
d=json.loads(l); print('$1', 'value', round(d['value']), 'ms', round(d['ms_per_step'],2))
PY
}
run default
BFB200_FUSED_PAIR=compact run compact
BFB200_FUSED_PAIR=compact timeout 900 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -3
BFB200_FUSED_PAIR=compact timeout 600 python tools/fused_accuracy.py 2>&1 | tail -6

#!/bin/bash
# quick check of a fused-kernel change: the fused parity tests, then the headline pass
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-aux --no-extras 2>&1 | tail -1 | cut -c1-260

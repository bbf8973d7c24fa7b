#!/bin/bash
# round-2 first GPU check of the rewritten fused kernel: parity tests, accuracy statistics, a short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2a_fused_tests.log 2>&1; echo "fused tests exit $?" 
tail -15 gpurun_out/r2a_fused_tests.log
timeout 600 python tools/fused_accuracy.py > gpurun_out/r2a_accuracy.log 2>&1; echo "accuracy exit $?"; cat gpurun_out/r2a_accuracy.log | tail -12
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/r2a_bench.log 2>&1; echo "bench exit $?"; tail -3 gpurun_out/r2a_bench.log

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -x -k "narrow_parity or full_width or banks" > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest model exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
HV_TRACE=gpurun_out/trace.csv timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
echo "== bench exit $?"; tail -n 1 gpurun_out/bench.log | cut -c1-300

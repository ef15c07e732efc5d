#!/bin/bash
mkdir -p gpurun_out
for pe in 0 4 3 2; do HV_ATTN_POLY=$pe timeout -s KILL 200 python scripts/attn_bench.py 48 6912 8 40 2>&1 | tail -1; done
for pe in 0 4 2; do HV_ATTN_POLY=$pe timeout -s KILL 200 python scripts/attn_bench.py 48 1728 8 80 2>&1 | tail -1; done
timeout -s KILL 300 python scripts/gemm_bst_ab.py > gpurun_out/gelu.log 2>&1; echo "== gemm exit $?"; cat gpurun_out/gelu.log | tail -10
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_ops.log 2>&1
echo "== pytest ops exit $?"; tail -3 gpurun_out/pytest_ops.log
HV_TRACE=gpurun_out/trace_14.txt timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_14.log 2>&1
echo "== bench exit $?"; tail -n 1 gpurun_out/bench_14.log | cut -c1-200; tail -n 1 gpurun_out/bench_14.log | grep -o '"op_profile.*' | cut -c1-900

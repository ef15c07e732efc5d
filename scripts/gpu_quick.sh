#!/bin/bash
# Quick GPU check: operator parity tests, GEMM / attention microbenchmarks, one bench line.
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_ops.log 2>&1
echo "== pytest ops exit $?"; tail -3 gpurun_out/pytest_ops.log
timeout -s KILL 300 python scripts/gemm_bst_ab.py > gpurun_out/gemm_ab.log 2>&1; echo "== gemm exit $?"; tail -9 gpurun_out/gemm_ab.log
timeout -s KILL 200 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -1
timeout -s KILL 200 python scripts/attn_bench.py 48 1728 8 80 2>&1 | tail -1
HV_TRACE=gpurun_out/trace_quick.txt timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
echo "== bench exit $?"; tail -n 1 gpurun_out/bench_quick.log | cut -c1-200; tail -n 1 gpurun_out/bench_quick.log | grep -o '"op_profile.*' | cut -c1-900

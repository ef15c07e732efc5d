#!/bin/bash
mkdir -p gpurun_out
HV_GEMM_BST=0 timeout -s KILL 300 python scripts/gemm_bst_ab.py > gpurun_out/bst0.log 2>&1; echo "== bst0 exit $?"; cat gpurun_out/bst0.log | tail -12
HV_GEMM_BST=1 timeout -s KILL 300 python scripts/gemm_bst_ab.py > gpurun_out/bst1.log 2>&1; echo "== bst1 exit $?"; cat gpurun_out/bst1.log | tail -12
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_ops.log 2>&1
echo "== pytest ops exit $?"; tail -3 gpurun_out/pytest_ops.log
HV_TRACE=gpurun_out/trace_bst.txt timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bst.log 2>&1
echo "== bench exit $?"; tail -n 1 gpurun_out/bench_bst.log | cut -c1-1500

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_ops.log 2>&1
echo "== pytest ops exit $?"; tail -5 gpurun_out/pytest_ops.log
HV_GEMM_TMA_IO=0 timeout -s KILL 300 python scripts/gemm_bst_ab.py > gpurun_out/io0.log 2>&1; echo "== io0 exit $?"; cat gpurun_out/io0.log | tail -10
HV_GEMM_TMA_IO=1 timeout -s KILL 300 python scripts/gemm_bst_ab.py > gpurun_out/io1.log 2>&1; echo "== io1 exit $?"; cat gpurun_out/io1.log | tail -10
HV_GEMM_TMA_IO=1 HV_GEMM_BST=1 timeout -s KILL 300 python scripts/gemm_bst_ab.py > gpurun_out/io1b.log 2>&1; echo "== io1+bst exit $?"; cat gpurun_out/io1b.log | tail -10
HV_TRACE=gpurun_out/trace_io.txt timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_io.log 2>&1
echo "== bench exit $?"; tail -n 1 gpurun_out/bench_io.log | cut -c1-200; tail -n 1 gpurun_out/bench_io.log | grep -o '"op_profile.*' | cut -c1-900
timeout -s KILL 900 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_model.log 2>&1
echo "== pytest model exit $?"; tail -5 gpurun_out/pytest_model.log

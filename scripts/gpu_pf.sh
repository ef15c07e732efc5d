#!/bin/bash
mkdir -p gpurun_out
for pf in 0 16 40; do echo "== HV_GEMM_PF=$pf"; HV_GEMM_PF=$pf timeout -s KILL 300 python scripts/gemm_bst_ab.py 2>&1 | tail -9; done
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x -k "gemm or conv" > gpurun_out/pytest_gemm.log 2>&1
echo "== pytest gemm/conv exit $?"; tail -3 gpurun_out/pytest_gemm.log
HV_TRACE=gpurun_out/trace_pf.txt timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pf.log 2>&1
echo "== bench exit $?"; tail -n 1 gpurun_out/bench_pf.log | cut -c1-200; tail -n 1 gpurun_out/bench_pf.log | grep -o '"op_profile.*' | cut -c1-900

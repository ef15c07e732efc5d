#!/bin/bash
mkdir -p gpurun_out
HV_ATTN_WARPS=16 timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention" -p no:cacheprovider > gpurun_out/ops_attention16.log 2>&1
echo "== ops attention (16 warps) exit $?"; tail -n 5 gpurun_out/ops_attention16.log
HV_ATTN_WARPS=16 HV_TRACE=gpurun_out/trace16.csv timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench16.log 2>&1
echo "== bench 16 exit $?"; tail -n 1 gpurun_out/bench16.log | cut -c1-200
HV_TRACE=gpurun_out/trace8.csv timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.log 2>&1
echo "== bench 8 exit $?"; tail -n 1 gpurun_out/bench8.log | cut -c1-200

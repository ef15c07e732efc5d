#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention" -p no:cacheprovider > gpurun_out/ops_attention8.log 2>&1
echo "== ops attention (8 warps) exit $?"; tail -n 3 gpurun_out/ops_attention8.log
HV_ATTN_WARPS=4 timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention" -p no:cacheprovider > gpurun_out/ops_attention4.log 2>&1
echo "== ops attention (4 warps) exit $?"; tail -n 3 gpurun_out/ops_attention4.log
HV_TRACE=gpurun_out/trace8.csv timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.log 2>&1
echo "== bench 8 exit $?"; tail -n 1 gpurun_out/bench8.log | cut -c1-200
HV_ATTN_WARPS=4 HV_TRACE=gpurun_out/trace4.csv timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.log 2>&1
echo "== bench 4 exit $?"; tail -n 1 gpurun_out/bench4.log | cut -c1-200

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 0 -c 14 -o gpurun_out/prof_gemm_v4 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
echo "== ncu gemm exit $?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel8 -s 0 -c 1 -o gpurun_out/prof_attn8 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
echo "== ncu attn exit $?"

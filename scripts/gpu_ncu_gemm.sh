#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm960_v2 python scripts/gemm_one.py 331776 960 320 0 0 > gpurun_out/ncu_gemm960.log 2>&1
echo "== ncu 960 exit $?"; tail -2 gpurun_out/ncu_gemm960.log

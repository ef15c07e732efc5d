#!/bin/bash
mkdir -p gpurun_out
HV_ATTN_POLY=0 timeout -s KILL 200 python scripts/attn_bench.py 48 6912 8 40 2>&1 | tail -3
HV_ATTN_POLY=4 timeout -s KILL 200 python scripts/attn_bench.py 48 6912 8 40 2>&1 | tail -3
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_geglu_L0 python scripts/gemm_one.py 331776 2560 320 1 0 > gpurun_out/ncu_geglu.log 2>&1
echo "== ncu geglu exit $?"; tail -2 gpurun_out/ncu_geglu.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm960_L0 python scripts/gemm_one.py 331776 960 320 0 0 > gpurun_out/ncu_gemm960.log 2>&1
echo "== ncu 960 exit $?"; tail -2 gpurun_out/ncu_gemm960.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemmres_L0 python scripts/gemm_one.py 331776 320 1280 0 1 > gpurun_out/ncu_gemmres.log 2>&1
echo "== ncu res exit $?"; tail -2 gpurun_out/ncu_gemmres.log

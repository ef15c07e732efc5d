#!/bin/bash
# Full ncu captures (--set full, source-level sampling) of the top kernels inside one bench step; summaries -> profiles/ with
#   python scripts/summarize_ncu.py gpurun_out/prof_<x>.ncu-rep profiles/r0N_ncu_<x>.txt
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 0 -c 14 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
echo "== ncu gemm exit $?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:attn_pp_kernel -s 0 -c 1 -o gpurun_out/prof_attn -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
echo "== ncu attn exit $?"
ls -la gpurun_out/*.ncu-rep

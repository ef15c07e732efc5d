#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 HV_ATTN_POLY=0 timeout -s KILL 200 python scripts/attn_bench.py 48 6912 8 40 > gpurun_out/attn_l0.log 2>&1; grep -v "^$" gpurun_out/attn_l0.log | tail -12
for pe in 0 11 12 13; do HV_ATTN_POLY=$pe timeout -s KILL 200 python scripts/attn_bench.py 48 1728 8 80 2>&1 | tail -1; done
for pe in 0 11 12 13; do HV_ATTN_POLY=$pe timeout -s KILL 200 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -1; done

#!/bin/bash
# ncu launch list of the bench command (per-launch gpu__time_duration, cold caches, serialised): profiles/ evidence for kernel shares
mkdir -p gpurun_out
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== ncu launch list exit $?"; tail -n 2 gpurun_out/ncu_bench.log | cut -c1-200; wc -l gpurun_out/launches.csv

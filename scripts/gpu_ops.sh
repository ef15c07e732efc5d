#!/bin/bash
# Runs the operator parity tests group by group, each in its own process under a hard timeout so a hung kernel in
# one group cannot take the others down.  Output -> gpurun_out/ops_<group>.log
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for g in gemm conv3x3 groupnorm layernorm temporal attention layout; do
  timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "$g" -p no:cacheprovider > gpurun_out/ops_$g.log 2>&1
  echo "== $g exit $?" | tee -a gpurun_out/ops_summary.log
  tail -n 25 gpurun_out/ops_$g.log
done

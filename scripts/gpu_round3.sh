#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest ops+model exit $?"; tail -n 8 gpurun_out/pytest_gpu.log
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
echo "== bench exit $?"; tail -n 2 gpurun_out/bench.log
K='regex:gemm_kernel|attn_kernel|temporal_attn|gn_|layernorm|small_linear|conv3x3_direct|upsample|nhwc|ncfhw|timestep'
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 1720 -c 860 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== ncu list exit $?"; wc -l gpurun_out/launches.csv

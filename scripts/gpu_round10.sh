#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python scripts/gemm_microbench.py > gpurun_out/gemm_microbench.log 2>&1
echo "== microbench exit $?"; cat gpurun_out/gemm_microbench.log | tail -16
timeout -s KILL 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest -m gpu exit $?"; grep -E "passed|failed|config2|config3|narrow:|full-width" gpurun_out/pytest_gpu.log | tail -12

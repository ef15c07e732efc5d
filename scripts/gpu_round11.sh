#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_pipeline_gpu.py tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -s -k "pipeline or temporal" > gpurun_out/pytest_pipe.log 2>&1
echo "== pytest pipeline+temporal exit $?"; grep -E "passed|failed|Error|pipeline 48|assert" gpurun_out/pytest_pipe.log | tail -8
timeout -s KILL 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.log 2>&1
echo "== bench (with cpu baseline) exit $?"; tail -n 1 gpurun_out/bench_full.log | cut -c1-300
timeout -s KILL 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
echo "== bench reference exit $?"; tail -n 1 gpurun_out/bench_ref.log | cut -c1-400
nproc; free -g | head -2

#!/bin/bash
# What the driver does at round end, plus a first bench line.
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest -m gpu exit $?"; tail -n 15 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "== smoke exit $?"; tail -n 3 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 ${BENCH_ARGS:---no-cpu-baseline} > gpurun_out/bench.log 2>&1
echo "== bench exit $?"; tail -n 5 gpurun_out/bench.log

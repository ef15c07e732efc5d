#!/bin/bash
# Full GPU check on one B200: parity suite, smoke(), bench (native arm with CPU baseline, reference arm).
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x -s > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest -m gpu exit $?"; grep -E "passed|failed|error|rel err|vs " gpurun_out/pytest_gpu.log | tail -12
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -3 gpurun_out/smoke.log
HV_TRACE=gpurun_out/trace_full.txt timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.log 2>&1
echo "== bench exit $?"; tail -n 1 gpurun_out/bench_full.log | cut -c1-400
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
echo "== bench reference exit $?"; tail -n 1 gpurun_out/bench_ref.log | cut -c1-300

#!/bin/bash
# Final round check: full parity suite, smoke(), bench config 2 (default) and config 3 (reference banks on).
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest -m gpu exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py > gpurun_out/bench_default.log 2>&1
echo "== bench (default flags) exit $?"; tail -n 1 gpurun_out/bench_default.log | cut -c1-260
timeout -s KILL 600 python bench.py --banks 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_banks.log 2>&1
echo "== bench --banks 1 exit $?"; tail -n 1 gpurun_out/bench_banks.log | cut -c1-260

#!/bin/bash
# Round-end check on one B200: guarded attention smoke, smoke(), full GPU suite, the bench lines of configs 2 / 3 / 5 and the CPU reference arm.
mkdir -p gpurun_out
timeout -s KILL 90 python scripts/attn_bench.py 12 6912 8 40 > gpurun_out/attn_gate.log 2>&1; rc=$?; tail -1 gpurun_out/attn_gate.log
if [ $rc -ne 0 ]; then echo "== attention gate FAILED rc=$rc: stopping"; exit 1; fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 500 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest rc=$?"; tail -26 gpurun_out/pytest_gpu.log
HV_TRACE=gpurun_out/trace_c2_final.csv timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_c2_final.log 2>&1; echo "== bench c2 rc=$?"; tail -n 1 gpurun_out/bench_c2_final.log | cut -c1-260
timeout 600 python bench.py --config 3 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_c3_final.log 2>&1; echo "== bench c3 rc=$?"; tail -n 1 gpurun_out/bench_c3_final.log | cut -c1-260
timeout 600 python bench.py --config 5 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_final.log 2>&1; echo "== bench c5 rc=$?"; tail -n 1 gpurun_out/bench_c5_final.log | cut -c1-260
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref_final.log 2>&1; echo "== bench reference rc=$?"; tail -n 1 gpurun_out/bench_ref_final.log | cut -c1-400

#!/bin/bash
# Last check of the round on one B200: guarded attention gate, smoke(), the full GPU suite, and `python bench.py` exactly as the driver runs it (wall time printed).
mkdir -p gpurun_out
timeout -s KILL 90 python scripts/attn_bench.py 12 6912 8 40 > gpurun_out/attn_gate.log 2>&1; rc=$?; tail -1 gpurun_out/attn_gate.log
if [ $rc -ne 0 ]; then echo "== attention gate FAILED rc=$rc: stopping"; exit 1; fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 500 > gpurun_out/pytest_gpu_last.log 2>&1; echo "== pytest rc=$?"; tail -24 gpurun_out/pytest_gpu_last.log
SECONDS=0
timeout 900 python bench.py > gpurun_out/bench_default_last.log 2>&1; echo "== default bench rc=$? wall ${SECONDS}s"; tail -n 1 gpurun_out/bench_default_last.log | cut -c1-260

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x -k attention > gpurun_out/pytest_attn.log 2>&1
echo "== pytest attention exit $?"; tail -3 gpurun_out/pytest_attn.log
HV_ATTN_POLY=31 timeout -s KILL 200 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -14
for pe in 0 13; do HV_ATTN_POLY=$pe timeout -s KILL 200 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -1; done
HV_ATTN_POLY=0 timeout -s KILL 200 python scripts/attn_bench.py 48 1728 8 80 2>&1 | tail -1
HV_ATTN_POLY=0 timeout -s KILL 200 python scripts/attn_bench.py 48 432 8 160 2>&1 | tail -1

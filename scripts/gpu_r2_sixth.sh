#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
HV_LIB=humanvid_b200/lib/libhv_b200_prev.so timeout -s KILL 200 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -1 | sed "s/^/prev /"
timeout -s KILL 200 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -1 | sed "s/^/new  /"
done | tee gpurun_out/attn_ab5.log
timeout -s KILL 200 python scripts/attn_bench.py 4 9216 8 40 2>&1 | tail -1
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest rc=$?"; tail -9 gpurun_out/pytest_gpu.log
timeout -s KILL 200 python scripts/pg_trace.py 2>&1 | tail -11
timeout -s KILL 200 python scripts/norm_bench.py 2>&1 | tail -3
HV_TRACE=gpurun_out/trace_c2_f.csv timeout 1200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager --no-extras > gpurun_out/bench_c2_f.log 2>&1; echo "== bench c2 rc=$?"; tail -n 1 gpurun_out/bench_c2_f.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['op_profile'], d['roofline']['traffic'])"

#!/bin/bash
# guarded: a 60 s attention check first; if it hangs or fails nothing else runs
mkdir -p gpurun_out
timeout -s KILL 60 python scripts/attn_bench.py 12 6912 8 40 > gpurun_out/attn_gate.log 2>&1; rc=$?; tail -1 gpurun_out/attn_gate.log
if [ $rc -ne 0 ]; then echo "== attention gate FAILED rc=$rc: stopping"; exit 1; fi
for i in 1 2; do
HV_LIB=humanvid_b200/lib/libhv_b200_prev.so timeout -s KILL 60 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -1 | sed "s/^/prev  (exchange)      /"
HV_LIB=humanvid_b200/lib/libhv_b200_prev2.so timeout -s KILL 60 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -1 | sed "s/^/prev2 (full-row max) /"
timeout -s KILL 60 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -1 | sed "s/^/new   (+ setmaxnreg) /"
done | tee gpurun_out/attn_ab5.log
timeout -s KILL 60 python scripts/attn_bench.py 4 9216 8 40 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x --timeout 400 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest rc=$?"; tail -24 gpurun_out/pytest_gpu.log
timeout -s KILL 100 python scripts/pg_trace.py 2>&1 | tail -11
HV_TRACE=gpurun_out/trace_c2_f.csv timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager --no-extras > gpurun_out/bench_c2_f.log 2>&1; echo "== bench c2 rc=$?"; tail -n 1 gpurun_out/bench_c2_f.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['op_profile'], d['roofline']['traffic'])"

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_fullsize_gpu.py -m gpu -q -p no:cacheprovider --timeout 400 -x > gpurun_out/pytest_part.log 2>&1; echo "== pytest rc=$?"; tail -6 gpurun_out/pytest_part.log | cut -c1-300
timeout -s KILL 120 python scripts/norm_bench.py 2>&1 | tail -5
HV_TRACE=gpurun_out/trace_c2_h.csv timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager --no-extras > gpurun_out/bench_c2_h.log 2>&1; echo "== bench c2 rc=$?"; tail -n 1 gpurun_out/bench_c2_h.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['achieved_tflops'], d['op_profile'])"
grep gemm_vt gpurun_out/trace_c2_h.csv | awk -F, '{a[$4" "$5" "$6]+=$7; n[$4" "$5" "$6]++} END{for(k in a) print "gemm_vt", k, n[k], a[k]/n[k]}'

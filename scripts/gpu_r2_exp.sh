#!/bin/bash
mkdir -p gpurun_out
T=humanvid_b200/lib/libhv_b200_tuning.so
for mb in 0 32 56 80; do HV_LIB=$T HV_GN_CHUNK_MB=$mb timeout -s KILL 120 python scripts/norm_bench.py 2>&1 | tail -5; done | tee gpurun_out/norm_ab.log
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_fullsize_gpu.py tests/test_pipeline_gpu.py -m gpu -q -p no:cacheprovider --timeout 400 -x > gpurun_out/pytest_part.log 2>&1; echo "== pytest rc=$?"; tail -8 gpurun_out/pytest_part.log | cut -c1-300
HV_TRACE=gpurun_out/trace_c2_g.csv timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager > gpurun_out/bench_c2_g.log 2>&1; echo "== bench c2 rc=$?"; tail -n 1 gpurun_out/bench_c2_g.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['achieved_tflops'], d['op_profile']); print(d.get('pipeline_clip'))"

#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout -s KILL 200 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -1
timeout -s KILL 200 python scripts/norm_bench.py 2>&1 | tail -3 | tee gpurun_out/norm_bench.log
timeout -s KILL 200 python scripts/pg_trace.py 2>&1 | tail -16
HV_TRACE=gpurun_out/trace_c2_d.csv timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-extras > gpurun_out/bench_c2_d.log 2>&1; echo "== bench rc=$?"; tail -n 1 gpurun_out/bench_c2_d.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['op_profile'])"
timeout -s KILL 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:gemm_kernel|attn_|gn_|layernorm|small_linear|nhwc|ncfhw|timestep|temporal' --launch-skip 1640 --launch-count 820 --csv --log-file gpurun_out/dram.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-extras > gpurun_out/ncu_dram.log 2>&1; echo "== ncu dram rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:attn_pp_kernel -s 0 -c 1 -o gpurun_out/prof_attn_pp_v2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-extras > gpurun_out/ncu_attn_pp2.log 2>&1; echo "== ncu attn rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 0 -c 48 -o gpurun_out/prof_gemm_v2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-extras > gpurun_out/ncu_gemm2.log 2>&1; echo "== ncu gemm rc=$?"
ls -la gpurun_out/*.ncu-rep

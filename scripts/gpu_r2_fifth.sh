#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout -s KILL 200 python scripts/pg_trace.py 2>&1 | tail -12
timeout -s KILL 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:gemm_kernel|attn_|gn_|layernorm|small_linear|nhwc|ncfhw|timestep|temporal' --launch-skip 1640 --launch-count 820 --csv --log-file gpurun_out/dram.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-extras > gpurun_out/ncu_dram.log 2>&1; echo "== ncu dram rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:attn_pp_kernel -s 0 -c 1 -o gpurun_out/prof_attn_pp_v2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-extras > gpurun_out/ncu_attn_pp2.log 2>&1; echo "== ncu attn rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 18 -o gpurun_out/prof_gemm_v2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-extras > gpurun_out/ncu_gemm2.log 2>&1; echo "== ncu gemm rc=$?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
HV_TRACE=gpurun_out/trace_c2_e.csv timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_final.log 2>&1; echo "== bench c2 rc=$?"; tail -n 1 gpurun_out/bench_c2_final.log | cut -c1-300
timeout -s KILL 400 python scripts/graph_gain.py 2>&1 | tail -4

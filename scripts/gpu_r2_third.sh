#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest rc=$?"; tail -22 gpurun_out/pytest_gpu.log
for m in 0 4 3 2; do HV_ATTN_POLY=$m timeout -s KILL 200 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -1; done | tee gpurun_out/attn_ab4.log
HV_ATTN_MW=1 timeout -s KILL 200 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -1 | sed "s/^/MW=1 /" | tee -a gpurun_out/attn_ab4.log
for k in 2816 1280; do HV_GEMM_MT2_MINK=$k timeout -s KILL 300 python scripts/gemm_ff_bench.py 2>&1 | tail -3; done | tee gpurun_out/gemm_ff.log
HV_TRACE=gpurun_out/trace_c2_c.csv timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.log 2>&1; echo "== bench c2 rc=$?"; tail -n 1 gpurun_out/bench_c2.log | cut -c1-400
timeout 900 python bench.py --steps 10 --warmup 3 --config 3 --no-cpu-baseline --no-extras > gpurun_out/bench_c3.log 2>&1; echo "== bench c3 rc=$?"; tail -n 1 gpurun_out/bench_c3.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 --config 5 --no-cpu-baseline > gpurun_out/bench_c5.log 2>&1; echo "== bench c5 rc=$?"; tail -n 1 gpurun_out/bench_c5.log | cut -c1-300
timeout -s KILL 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --launch-skip 1700 --launch-count 1700 --csv --log-file gpurun_out/dram.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-extras > gpurun_out/ncu_dram.log 2>&1; echo "== ncu dram rc=$?"

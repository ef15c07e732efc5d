#!/bin/bash
# Round-2 first GPU pass: smoke, full GPU test suite, per-layer error ladder, bench (config 2) with trace, ncu of the shipped L0 attention kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
timeout 900 python scripts/error_ladder.py --out gpurun_out/error_ladder_config2.txt > gpurun_out/ladder.log 2>&1; echo "== ladder rc=$?"; tail -5 gpurun_out/ladder.log
HV_TRACE=gpurun_out/trace_c2.csv timeout 900 python bench.py --steps 10 --warmup 3 --quick-cpu > gpurun_out/bench_c2.log 2>&1; echo "== bench rc=$?"; tail -c 3000 gpurun_out/bench_c2.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:attn_pp_kernel -s 0 -c 1 -o gpurun_out/prof_attn_pp -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager > gpurun_out/ncu_attn_pp.log 2>&1
echo "== ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep
tools/mufu_rate > gpurun_out/mufu_rate.log 2>&1; cat gpurun_out/mufu_rate.log
for m in 0 16 17; do HV_ATTN_POLY=$m timeout -s KILL 200 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -1; done | tee gpurun_out/attn_ab.log
HV_ATTN_POLY=31 timeout -s KILL 200 python scripts/attn_bench.py 12 6912 8 40 > gpurun_out/attn_phases.log 2>&1; tail -12 gpurun_out/attn_phases.log

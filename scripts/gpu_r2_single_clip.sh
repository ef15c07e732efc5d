#!/bin/bash
# Strong scaling of ONE clip over 2 GPUs: config 3 split into its two CFG halves and config 5's six (window x CFG-half) units.
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 2 --config 3 --single-clip --steps 10 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/bench_c3_single_clip_2gpu.log 2>&1
echo "== config 3 single clip rc=$?"; tail -n 1 gpurun_out/bench_c3_single_clip_2gpu.log | cut -c1-300
timeout 600 python bench.py --gpus 2 --config 5 --steps 6 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/bench_c5_2gpu_b.log 2>&1
echo "== config 5 rc=$?"; tail -n 1 gpurun_out/bench_c5_2gpu_b.log | cut -c1-300
grep -o '"solo_rank0": {[^}]*}' gpurun_out/bench_c3_single_clip_2gpu.log gpurun_out/bench_c5_2gpu_b.log

#!/bin/bash
# 2-GPU check of the cost-balanced unit assignment: bitwise dist_check, the multi-GPU test, config 5 with its in-run single-GPU reference.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/dist_check.py > gpurun_out/dist_check_b_$N.log 2>&1; echo "== dist_check rc=$?"; grep dist_check gpurun_out/dist_check_b_$N.log | tail -2
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --gpus $N --config 5 --steps 6 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/bench_c5_${N}gpu_c.log 2>&1; echo "== bench config5 x$N rc=$?"; tail -n 1 gpurun_out/bench_c5_${N}gpu_c.log | cut -c1-200
grep -o '"solo_rank0": {[^}]*}' gpurun_out/bench_c5_${N}gpu_c.log

#!/bin/bash
# 8-GPU pass (gpurun --gpus 8, charged 8x: kept short): unit split check, config 4 (8 clips) and config 5 (one clip, 6 units on 8 ranks)
N=${1:-8}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/dist_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "== dist_check rc=$?"; grep "dist_check world" gpurun_out/dist_check_$N.log | tail -1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/bench_c4_$N.log 2>&1; echo "== bench config4 x$N rc=$?"; tail -n 1 gpurun_out/bench_c4_$N.log | cut -c1-330
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --config 5 --steps 8 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/bench_c5_$N.log 2>&1; echo "== bench config5 x$N rc=$?"; tail -n 1 gpurun_out/bench_c5_$N.log | cut -c1-330

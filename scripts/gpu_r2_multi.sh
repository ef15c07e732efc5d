#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): NCCL unit split vs single-GPU loop, config 4 (N clips) and config 5 (one clip, 6 units) scaling lines.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/dist_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "== dist_check rc=$?"; grep dist_check gpurun_out/dist_check_$N.log | tail -2; tail -3 gpurun_out/dist_check_$N.log
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_$N.log 2>&1; echo "== bench config4 x$N rc=$?"; tail -n 1 gpurun_out/bench_c4_$N.log | cut -c1-330
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_$N.log 2>&1; echo "== bench config5 x$N rc=$?"; tail -n 1 gpurun_out/bench_c5_$N.log | cut -c1-330

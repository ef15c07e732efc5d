#!/bin/bash
mkdir -p gpurun_out
for g in gemm conv3x3 layernorm; do
  timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "$g" -p no:cacheprovider > gpurun_out/ops_$g.log 2>&1
  echo "== ops $g exit $?"; tail -n 3 gpurun_out/ops_$g.log
done
timeout -s KILL 600 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -x -k "narrow_parity or full_width or banks or golden" > gpurun_out/pytest_model.log 2>&1
echo "== pytest model exit $?"; tail -n 4 gpurun_out/pytest_model.log
HV_TRACE=gpurun_out/trace.csv timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
echo "== bench exit $?"; tail -n 1 gpurun_out/bench.log | cut -c1-200

#!/bin/bash
# Validation of the mma.sync PoseGuider conv_in (smallconv.cu::pg_conv_in_kernel): operator parity incl. ragged tiles, every PoseGuider / pipeline
# test that runs it, per-launch device times, smoke(); then ncu --set full of the kernel and of the level-1/2 attention / temporal kernels.
mkdir -p gpurun_out
SECONDS=0
timeout 150 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_config5_gpu.py -m gpu -q -x -p no:cacheprovider --timeout 120 \
  -k "pose_conv_in or conv3x3_small or conv3x3_direct or pose_guider" > gpurun_out/pytest_pg.log 2>&1; echo "== pytest pg rc=$? at ${SECONDS}s"; tail -4 gpurun_out/pytest_pg.log
timeout 60 python scripts/pg_trace.py > gpurun_out/pg_trace.log 2>&1; echo "== pg_trace rc=$? at ${SECONDS}s"; head -14 gpurun_out/pg_trace.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== smoke done at ${SECONDS}s"
timeout -s KILL 60 ncu --set full --clock-control none --import-source on -k 'regex:pg_conv_in' -s 2 -c 1 \
  -o gpurun_out/prof_pg_conv_in -f python scripts/cond_once.py > gpurun_out/ncu_pg.log 2>&1
echo "== ncu pg exit $? at ${SECONDS}s"
timeout -s KILL 100 ncu --set full --clock-control none --import-source on -k 'regex:attn_kernel8|temporal_attn_kernel' -s 0 -c 8 \
  -o gpurun_out/prof_attn8_tattn -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-eager --no-extras > gpurun_out/ncu_attn8.log 2>&1
echo "== ncu attn8 exit $? at ${SECONDS}s"
ls -la gpurun_out/*.ncu-rep

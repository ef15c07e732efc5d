#!/bin/bash
# ncu --set full captures of the kernels that had no committed capture yet (north_star: "each kernel evidenced by a committed ncu capture"):
# GroupNorm stats / apply, LayerNorm (plain, +cross-attention vector, +positional encoding), temporal attention at level 0, and the
# conditioning branches' small-channel / Pluecker / PixelUnshuffle kernels; then smoke() and the GPU suite in whatever time is left.
# Summaries -> profiles/ with scripts/summarize_ncu.py.
mkdir -p gpurun_out
SECONDS=0
# first 16 matching launches of a forward = down block 0: resnet GN1/GN2 (stats, apply), transformer GN, LN1, LN3(+attn2 vector), motion-module GN,
# LN+PE, temporal attention (d = 40, TMA kernel), LN+PE, temporal attention, ff_norm LN
timeout -s KILL 260 ncu --set full --clock-control none --import-source on -k 'regex:gn_stats|gn_apply|layernorm_kernel|temporal_attn' -s 0 -c 16 \
  -o gpurun_out/prof_norm_tattn -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-eager --no-extras > gpurun_out/ncu_norm_tattn.log 2>&1
echo "== ncu norms/temporal exit $? at ${SECONDS}s"
timeout -s KILL 150 ncu --set full --clock-control none --import-source on -k 'regex:smallconv_mma|pg_conv_in|plucker_unshuffle|conv3x3_direct' -s 0 -c 8 \
  -o gpurun_out/prof_cond -f python scripts/cond_once.py > gpurun_out/ncu_cond.log 2>&1
echo "== ncu cond exit $? at ${SECONDS}s"; tail -n 1 gpurun_out/ncu_cond.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== smoke done at ${SECONDS}s"
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu_last.log 2>&1; echo "== pytest rc=$? at ${SECONDS}s"; tail -5 gpurun_out/pytest_gpu_last.log

#!/bin/bash
# Tile-shape A/B for the level-1 / level-2 linears on the tuning build (humanvid_b200/lib/libhv_b200_tuning.so = HV_BUILD_TUNING=1 build).
mkdir -p gpurun_out
export HV_LIB=$PWD/humanvid_b200/lib/libhv_b200_tuning.so
timeout 120 python scripts/gemm_tile_ab.py > gpurun_out/tile_ab_default.log 2>&1; tail -n 14 gpurun_out/tile_ab_default.log
HV_GEMM_BN=128 timeout 120 python scripts/gemm_tile_ab.py > gpurun_out/tile_ab_bn128.log 2>&1; tail -n 14 gpurun_out/tile_ab_bn128.log
HV_GEMM_BN=128 HV_GEMM_MT2_MINK=512 timeout 120 python scripts/gemm_tile_ab.py > gpurun_out/tile_ab_bn128_mt2.log 2>&1; tail -n 14 gpurun_out/tile_ab_bn128_mt2.log
HV_GEMM_MT2_MINK=512 timeout 120 python scripts/gemm_tile_ab.py > gpurun_out/tile_ab_mt2.log 2>&1; tail -n 14 gpurun_out/tile_ab_mt2.log

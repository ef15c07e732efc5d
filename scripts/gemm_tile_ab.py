"""A/B of the linear-GEMM tile choice on the level-1 / level-2 shapes (tuning build, HV_LIB=<tuning lib>):
    python scripts/gemm_tile_ab.py                      # release choice (128 x 160 / 128 x 256 tiles)
    HV_GEMM_BN=128 python scripts/gemm_tile_ab.py       # 128 x 128 tiles
    HV_GEMM_BN=128 HV_GEMM_MT2_MINK=512 python scripts/gemm_tile_ab.py   # 256 x 128 tiles, accumulators still double-buffered (2 x 2 x 128 = 512 TMEM columns)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_microbench import bench

SHAPES = [  # M, N, K, residual
    (82944, 640, 640, True), (82944, 640, 640, False), (82944, 640, 2560, True), (82944, 1920, 640, False), (82944, 1280, 640, False),
    (20736, 1280, 1280, True), (20736, 1280, 1280, False), (20736, 1280, 5120, True), (20736, 3840, 1280, False), (20736, 2560, 1280, False),
    (5184, 1280, 1280, True), (5184, 3840, 1280, False), (331776, 320, 1280, True),
]
tag = f"BN={os.environ.get('HV_GEMM_BN', 'auto')} MT2_MINK={os.environ.get('HV_GEMM_MT2_MINK', 'default')}"
tot = 0.0
for M, N, K, res in SHAPES:
    t = bench(M, N, K, res=res, iters=30)
    tot += t
    print(f"{tag}: M={M} N={N} K={K} res={int(res)}: {t:.4f} ms  {2 * M * N * K / t / 1e9:.0f} TF/s", flush=True)
print(f"{tag}: sum {tot:.3f} ms")

"""A/B of the B-stationary GEMM variant (HV_GEMM_BST=0/1, read once per process) on the linear shapes of levels 0/1."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_microbench import bench  # noqa: E402  (runs its own table on import only when executed as __main__)

if __name__ == "__main__":
    print("HV_GEMM_BST =", os.environ.get("HV_GEMM_BST", "(default 1)"))
    for (M, N, K, geglu, res) in [(331776, 320, 320, False, True), (331776, 960, 320, False, False), (331776, 320, 1280, False, True),
                                  (331776, 2560, 320, True, False), (82944, 640, 640, False, True), (82944, 1920, 640, False, False),
                                  (82944, 5120, 640, True, False), (82944, 640, 2560, False, True), (20736, 1280, 1280, False, True)]:
        t = bench(M, N, K, res=res, geglu=geglu)
        fl = 2.0 * M * N * K / t / 1e9
        print(f"M={M} N={N} K={K} geglu={int(geglu)} res={int(res)}: {t:.3f} ms  {fl:.0f} TF/s", flush=True)

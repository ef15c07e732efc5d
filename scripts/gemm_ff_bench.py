"""Timing of the feed-forward GEMMs (GEGLU FF1, FF2 + residual) at the three resolutions; env HV_GEMM_MT2_MINK A/B."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_microbench import bench
for (M, C) in [(331776, 320), (82944, 640), (20736, 1280)]:
    t1 = bench(M, 8 * C, C, geglu=True)
    t2 = bench(M, C, 4 * C)
    t3 = bench(M, 3 * C, C, res=False, bias=False)
    print(f"M={M} C={C} MT2_MINK={os.environ.get('HV_GEMM_MT2_MINK','default')}: FF1 geglu {t1:.3f} ms ({2*M*8*C*C/t1/1e9:.0f} TF/s)  FF2+res {t2:.3f} ms ({2*M*4*C*C/t2/1e9:.0f} TF/s)  QKV {t3:.3f} ms ({2*M*3*C*C/t3/1e9:.0f} TF/s)", flush=True)

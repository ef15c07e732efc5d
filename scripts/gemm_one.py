"""One GEMM shape a few times (for ncu): python scripts/gemm_one.py M N K geglu res"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_microbench import bench
M, N, K, geglu, res = [int(v) for v in sys.argv[1:6]]
print(bench(M, N, K, res=bool(res), geglu=bool(geglu), iters=3))

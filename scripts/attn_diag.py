"""Run one attention case in its own process: python scripts/attn_diag.py NF L heads d [Lb Fr nf_nobank]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_ops_gpu import attention_case
a = [int(v) for v in sys.argv[1:]]
kw = {}
if len(a) > 4:
    kw = dict(Lb=a[4], Fr=a[5], nf_nobank=a[6])
try:
    r = attention_case(a[0], a[1], a[2], a[3], **kw)
    print("CASE", a, "rel", r, flush=True)
except Exception as e:
    print("CASE", a, "FAILED", str(e).splitlines()[0], flush=True)

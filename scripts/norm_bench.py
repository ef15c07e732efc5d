"""Timing of hv_op_layernorm / hv_op_groupnorm at the level-0/1 shapes (HBM-bound passes): GB/s on read + write bytes."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanvid_b200._native as _N
if os.environ.get("HV_LIB"):   # A/B against another build of the library (e.g. the tuning build)
    _N.LIB_PATH = os.environ["HV_LIB"]
from humanvid_b200._native import check, i32, i64, lib, ptr, stream

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for (NF, HW, Cc) in [(48, 6912, 320), (48, 6912, 640), (48, 1728, 640), (48, 1728, 1280), (48, 432, 1280)]:
    rows = NF * HW
    x = torch.randn(rows, Cc, device="cuda").half()
    g, b = torch.ones(Cc, device="cuda").half(), torch.zeros(Cc, device="cuda").half()
    out = torch.empty_like(x)
    t_ln = timeit(lambda: check(lib().hv_op_layernorm(ptr(x), ptr(g), ptr(b), ptr(out), i64(rows), i64(Cc), C.c_float(1e-5), None, i64(1), None, None, i64(1), i64(1), stream())))
    f = lib().hv_groupnorm_scratch_floats
    f.restype = C.c_size_t
    stats = torch.zeros(int(f(i64(Cc), i64(NF), i64(HW), i32(32))), device="cuda", dtype=torch.float32)
    t_gn = timeit(lambda: check(lib().hv_op_groupnorm(ptr(x), i64(Cc), None, i64(0), ptr(g), ptr(b), ptr(out), i64(NF), i64(HW), i32(32), C.c_float(1e-5), i32(1), ptr(stats), stream())))
    nb = rows * Cc * 2
    print(f"rows={rows} C={Cc} GN_CHUNK_MB={os.environ.get('HV_GN_CHUNK_MB','56')}: layernorm {t_ln*1e3:.1f} us ({2*nb/t_ln/1e6:.0f} GB/s)   groupnorm+silu {t_gn*1e3:.1f} us ({3*nb/t_gn/1e6:.0f} GB/s on 2 reads + 1 write)", flush=True)

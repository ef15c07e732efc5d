"""Timing + accuracy of hv_op_attention at one shape: python scripts/attn_bench.py NF L heads d   (env HV_ATTN_POLY / HV_ATTN_WARPS)"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanvid_b200._native as _N
if os.environ.get("HV_LIB"):   # A/B against another build of the library
    _N.LIB_PATH = os.environ["HV_LIB"]
from humanvid_b200._native import check, i32, i64, lib, ptr, stream

NF, L, heads, d = [int(v) for v in sys.argv[1:5]]
dpad, dv, Lp = (d + 15) // 16 * 16, (d + 1 + 15) // 16 * 16, (L + 7) // 8 * 8
g = torch.Generator(device="cuda").manual_seed(1)
q, k, v = [torch.randn(NF * L, heads, d, generator=g, device="cuda").half() for _ in range(3)]
qk = torch.zeros(NF * L, 2, heads, dpad, device="cuda", dtype=torch.half)
qk[:, 0, :, :d], qk[:, 1, :, :d] = q, k
qk = qk.reshape(NF * L, 2 * heads * dpad)
vt = torch.zeros(heads, dv, NF, Lp, device="cuda", dtype=torch.half)
vt[:, :d, :, :L] = v.reshape(NF, L, heads, d).permute(2, 3, 0, 1)
vt[:, d, :, :L] = 1
out = torch.zeros(NF * L, heads * d, device="cuda", dtype=torch.half)
k_view = qk[:, heads * dpad:]
def run():
    check(lib().hv_op_attention(ptr(qk), C.c_void_p(k_view.data_ptr()), ptr(vt), ptr(out), i64(NF), i64(L), i32(heads), i32(d),
                                i64(2 * heads * dpad), i64(2 * heads * dpad), i64(NF * Lp), i64(heads * d), None, None, i64(0),
                                i64(heads * dpad), i64(0), i64(1), i64(0), i64(Lp), i64(0), stream()))
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
# accuracy on frame 0 against fp32 SDPA
qf, kf, vf = [t[:L].float().transpose(0, 1) for t in (q, k, v)]
ref = torch.nn.functional.scaled_dot_product_attention(qf, kf, vf).transpose(0, 1).reshape(L, heads * d)
err = float((out[:L].float() - ref).norm() / ref.norm())
print(f"attn NF={NF} L={L} heads={heads} d={d} poly={os.environ.get('HV_ATTN_POLY','default')} warps={os.environ.get('HV_ATTN_WARPS','8')}: {ms:.3f} ms  rel err {err:.2e}", flush=True)

"""Isolated timing of hv_op_gemm at the L0 linear shapes, with epilogue pieces switched off one by one."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanvid_b200._native as _N
if os.environ.get("HV_LIB"):   # A/B against another build of the library (e.g. the tuning build)
    _N.LIB_PATH = os.environ["HV_LIB"]
from humanvid_b200._native import Epilogue, check, i64, lib, ptr, stream

def bench(M, N, K, bias=True, res=True, n_valid=0, iters=20, geglu=False):
    A = torch.randn(M, K, device="cuda").half()
    W = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    b = torch.randn(N, device="cuda").half()
    R = torch.randn(M, N // 2 if geglu else N, device="cuda").half()
    out = torch.empty(M, N // 2 if geglu else N, device="cuda", dtype=torch.half)
    ep = Epilogue(bias=ptr(b).value if bias else None, residual=ptr(R).value if (res and not geglu) else None, ldr=N, geglu=int(geglu), n_valid=n_valid)
    def run():
        check(lib().hv_op_gemm(ptr(A), i64(K), None, i64(0), i64(0), ptr(W), ptr(out), i64(out.shape[1]), i64(M), i64(N), i64(K), C.byref(ep), stream()))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def main():
    M = 331776
    for (N, K) in [(320, 320), (960, 320), (320, 1280)]:
        for bn in ("", "128", "160", "256"):
            if bn: os.environ["HV_GEMM_BN"] = bn
            else: os.environ.pop("HV_GEMM_BN", None)
            t_full = bench(M, N, K)
            t_nores = bench(M, N, K, res=False)
            t_nobias = bench(M, N, K, bias=False, res=False)
            t_nostore = bench(M, N, K, bias=False, res=False, n_valid=8)
            print(f"N={N} K={K} BN={bn or 'auto'}: full {t_full:.3f}  no-res {t_nores:.3f}  no-bias/res {t_nobias:.3f}  no-store {t_nostore:.3f} ms", flush=True)
    os.environ.pop("HV_GEMM_BN", None)
    print("geglu N=2560 K=320:", round(bench(M, 2560, 320, geglu=True), 3), " plain N=2560 K=320 no-res:", round(bench(M, 2560, 320, res=False), 3), flush=True)


if __name__ == '__main__':
    main()

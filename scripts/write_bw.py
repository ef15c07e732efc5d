"""HBM ceilings that bound the wide-output K = 320 linears (QKV, GEGLU at level 0): pure-write, pure-read and copy bandwidth of this B200,
measured with torch's own kernels (fill_, sum, copy_) on buffers larger than L2.  Usage: python scripts/write_bw.py"""
import torch


def timed(fn, n=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    gb = 2
    x = torch.empty(gb << 30, dtype=torch.uint8, device="cuda")
    y = torch.empty(gb << 30, dtype=torch.uint8, device="cuda")
    xf = x.view(torch.float32)
    t = timed(lambda: x.fill_(1))
    print(f"write only (fill_ {gb} GiB):      {gb * 1.0737 / t * 1e3:7.0f} GB/s")
    t = timed(lambda: x.zero_())
    print(f"write only (zero_ {gb} GiB):      {gb * 1.0737 / t * 1e3:7.0f} GB/s")
    t = timed(lambda: xf.sum())
    print(f"read only  (sum {gb} GiB fp32):   {gb * 1.0737 / t * 1e3:7.0f} GB/s")
    t = timed(lambda: y.copy_(x))
    print(f"copy       ({gb} GiB -> {gb} GiB):   {2 * gb * 1.0737 / t * 1e3:7.0f} GB/s (read + write)")
    # 3 : 1 write : read mix of the level-0 QKV GEMM (A 212 MB in, 637 MB out): y[:3n] = f(x[:n])
    n = 1 << 28
    xs, yd = x[:n].view(torch.float16), y[: 3 * n].view(torch.float16).view(3, -1)
    t = timed(lambda: torch.mul(xs.unsqueeze(0), 2.0, out=yd) if False else yd.copy_(xs.unsqueeze(0).expand(3, -1)))
    print(f"1 read : 3 write (broadcast copy): {4 * n / 1e6 / t:7.0f} GB/s total, {3 * n / 1e6 / t:7.0f} GB/s of writes")


if __name__ == "__main__":
    main()

"""Per-shape table of the tcgen05 GEMM / conv launches of one forward, from the per-launch trace `HV_TRACE=<csv> python bench.py` writes:
launches, device time, TF/s, compulsory HBM GB/s (A once, residual once, output once, fp16), and each shape's own roofline bound
max(FLOPs / 1421.9 TF/s, reads / 6586 GB/s + writes / 3924 GB/s) with the fraction of it that is reached, sorted by time.
Usage: python scripts/gemm_shape_table.py gpurun_out/trace_c2_final.csv > profiles/r02_gemm_shapes.txt"""
import collections
import csv
import sys

TENSOR, HBM_R, HBM_W = 1421.9, 6586.4, 3924.0   # MEASURED_PEAKS.json (sustained bf16, copy bandwidth); profiles/r02_hbm_ceilings.txt (write only)


def work(lb, M, N, K):
    """-> flops, bytes read, bytes written"""
    if lb == "gemm_vt":
        return 2 * M * N * K, 2 * N * K, 2 * M * N
    if lb.startswith("conv3") or lb.startswith("upconv"):
        taps = 4 if lb.startswith("upconv") else 9
        rows_in = M / 4 if lb.startswith("upconv") else (M * 4 if lb.endswith("_s2") else M)
        return 2 * M * N * K, 2 * (rows_in * K / taps + (M * N if lb.endswith("_res") else 0)), 2 * M * N
    n_out = N / 2 if lb == "gemm_geglu" else N
    return 2 * M * N * K, 2 * (M * K + (M * n_out if lb == "gemm_res" else 0)), 2 * M * n_out


def main(path):
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, ""])
    for r in csv.DictReader(open(path)):
        if int(r["cat"]) not in (0, 1) or not r["label"]:
            continue
        M, N, K, ms = float(r["M"]), float(r["N"]), float(r["K"]), float(r["ms"])
        fl, rd, wr = work(r["label"], M, N, K)
        tt, th = fl / (TENSOR * 1e9), rd / (HBM_R * 1e6) + wr / (HBM_W * 1e6)
        a = agg[(r["label"], int(M), int(N), int(K))]
        a[0] += 1; a[1] += ms; a[2] += fl; a[3] += rd + wr; a[4] += max(tt, th); a[5] = "hbm" if th > tt else "tensor"
    tot, bound = sum(a[1] for a in agg.values()), sum(a[4] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.2f} ms, {sum(a[2] for a in agg.values()) / 1e9 / tot:.0f} TF/s; "
          f"sum of per-shape roofline bounds {bound:.2f} ms = {bound / tot:.3f} of the measured time")
    print(f"{'kind':<12}{'M':>8}{'N':>8}{'K':>7}{'n':>4}{'ms':>9}{'TF/s':>7}{'GB/s':>7}{'bound':>8}{'bound ms':>10}{'frac':>7}")
    for (lb, M, N, K), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{lb:<12}{M:>8}{N:>8}{K:>7}{a[0]:>4}{a[1]:>9.3f}{a[2] / 1e9 / a[1]:>7.0f}{a[3] / 1e6 / a[1]:>7.0f}{a[5]:>8}{a[4]:>10.3f}{a[4] / a[1]:>7.2f}")


if __name__ == "__main__":
    main(sys.argv[1])

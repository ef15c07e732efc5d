"""Per-shape table of the tcgen05 GEMM / conv launches of one forward, from the per-launch trace `HV_TRACE=<csv> python bench.py` writes:
launches, device time, TF/s, compulsory HBM GB/s (A once, output once, residual once, fp16) and arithmetic intensity, sorted by time.
Usage: python scripts/gemm_shape_table.py gpurun_out/trace_c2_final.csv > profiles/r02_gemm_shapes.txt"""
import collections
import csv
import sys


def flops_bytes(lb, M, N, K):
    if lb == "gemm_vt":
        return 2 * M * N * K, 2 * (N * K + M * N)
    if lb.startswith("conv3") or lb.startswith("upconv"):
        taps = 4 if lb.startswith("upconv") else 9
        rows_in = M / 4 if lb.startswith("upconv") else (M * 4 if lb.endswith("_s2") else M)
        return 2 * M * N * K, 2 * (rows_in * K / taps + M * N * (2 if lb.endswith("_res") else 1))
    n_out = N / 2 if lb == "gemm_geglu" else N
    return 2 * M * N * K, 2 * (M * K + M * n_out * (2 if lb == "gemm_res" else 1))


def main(path):
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for r in csv.DictReader(open(path)):
        if int(r["cat"]) not in (0, 1) or not r["label"]:
            continue
        M, N, K, ms = float(r["M"]), float(r["N"]), float(r["K"]), float(r["ms"])
        fl, by = flops_bytes(r["label"], M, N, K)
        a = agg[(r["label"], int(M), int(N), int(K))]
        a[0] += 1; a[1] += ms; a[2] += fl; a[3] += by
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.2f} ms, {sum(a[2] for a in agg.values()) / 1e9 / tot:.0f} TF/s")
    print(f"{'kind':<12}{'M':>8}{'N':>8}{'K':>7}{'n':>4}{'ms':>9}{'TF/s':>7}{'GB/s':>7}{'FLOP/B':>8}")
    for (lb, M, N, K), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{lb:<12}{M:>8}{N:>8}{K:>7}{a[0]:>4}{a[1]:>9.3f}{a[2] / 1e9 / a[1]:>7.0f}{a[3] / 1e6 / a[1]:>7.0f}{a[2] / a[3]:>8.0f}")


if __name__ == "__main__":
    main(sys.argv[1])

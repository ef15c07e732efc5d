"""Summarise an ncu launch list (gpu__time_duration.sum per launch, --csv) into per-kernel totals for ONE bench step.

    python scripts/summarize_launches.py gpurun_out/launches.csv profiles/r01_ncu_launches_step.csv profiles/r01_ncu_launches_summary.txt

The list covers every forward bench.py runs (warm-up, timed, e2e, profiling pass); the last complete forward is cut out by
looking for the step's first kernel (ncfhw_to_nhwc of the latents)."""
import csv, re, sys, collections

src, out_csv, out_txt = sys.argv[1:4]
rows = []
with open(src, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((int(r["ID"]), r["Kernel Name"], ns))
ours = re.compile(r"gemm_kernel|attn_|temporal_attn|gn_|groupnorm|layernorm|small_linear|upsample|nhwc|ncfhw|timestep|add_kernel|pixel_unshuffle|conv3x3_direct")
mine = [(i, k, ns) for (i, k, ns) in rows if ours.search(k)]
starts = [n for n, (i, k, ns) in enumerate(mine) if "ncfhw_to_nhwc" in k]
# one forward = from one latent-layout kernel to the next; take the last complete one
fw = None
for a, b in zip(starts[::-1][1:], starts[::-1][:-1]):
    if b - a > 400:
        fw = mine[a:b]
        break
if fw is None:
    fw = mine
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["ncu_id", "kernel", "gpu__time_duration_ns"])
    for i, k, ns in fw:
        w.writerow([i, re.sub(r"\(.*", "", k)[:90], int(ns)])
fam = collections.OrderedDict([("tcgen05 GEMM / implicit-GEMM conv (gemm_kernel)", "gemm_kernel"), ("temporal attention", "temporal_attn"),
                               ("spatial attention (attn_pp / attn_kernel8)", "attn_pp|attn_kernel"), ("GroupNorm", "gn_|groupnorm"), ("LayerNorm", "layernorm"), ("other (layout, upsample, small linear, ...)", ".")])
tot = collections.OrderedDict((k, [0, 0.0]) for k in fam)
for i, k, ns in fw:
    for name, pat in fam.items():
        if re.search(pat, k):
            tot[name][0] += 1
            tot[name][1] += ns
            break
total = sum(v[1] for v in tot.values())
with open(out_txt, "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 1 --no-cpu-baseline\n")
    f.write(f"# one UNet forward (config 2): {len(fw)} launches of this library's kernels, {total / 1e6:.2f} ms summed (cold-cache, serialised)\n")
    for name, (n, ns) in tot.items():
        f.write(f"{name:60s} {n:5d} launches {ns / 1e6:9.3f} ms  {100 * ns / total:5.1f} %\n")
print(open(out_txt).read())

"""ncu metrics pass (dram bytes + duration per launch) of one bench step -> profiles/r02_dram_traffic.{json,txt}.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'hv::' \
        --csv --log-file gpurun_out/dram.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-extras
    python scripts/summarize_dram.py gpurun_out/dram.csv profiles/r02_dram_traffic [launch_list.csv]

Every launch of the library in that process is captured; the summary covers the LAST complete UNet forward (from the last
`timestep_embedding` launch to the end).  With a third argument the per-launch list of that forward is written too.
"""
import collections
import csv
import json
import re
import sys


def family(name):
    for k in ("gemm_kernel", "attn_pp_kernel", "attn_kernel8", "temporal_attn", "gn_stats", "gn_finalize", "gn_apply", "layernorm", "small_linear",
              "ncfhw_to_nhwc", "nhwc_to_ncfhw", "timestep"):
        if k in name:
            return k
    return re.sub(r"<.*", "", name.split("(")[0]).strip()[-40:]


def main(src, dst, listing=None):
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    col = {h: i for i, h in enumerate(hdr)}
    per = collections.defaultdict(lambda: collections.defaultdict(float))   # launch id -> metric -> value
    names = {}
    for r in rd:
        if len(r) < len(hdr):
            continue
        lid = r[col["ID"]]
        names[lid] = r[col["Kernel Name"]]
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        m = r[col["Metric Name"]]
        if "bytes" in m:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        if "duration" in m:
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1)   # -> ms
        per[lid][m] = v
    ids = sorted(per, key=int)
    starts = [i for i in ids if "timestep_embedding" in names[i]]
    if starts:                                    # the last complete forward
        ids = [i for i in ids if int(i) >= int(starts[-1])]
    per = {i: per[i] for i in ids}
    if listing:
        with open(listing, "w") as f:
            f.write("# one UNet forward of `bench.py --steps 1` (config 2): ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                    "--clock-control none (cold-cache serialised replays)\nid,kernel,time_us,dram_read_MB,dram_write_MB\n")
            for n, i in enumerate(ids):
                m = per[i]
                f.write(f"{n},{re.sub(r'^.*hv::', '', names[i].split('(')[0])[:60]},{m.get('gpu__time_duration.sum', 0) * 1e3:.2f},"
                        f"{m.get('dram__bytes_read.sum', 0) / 1e6:.1f},{m.get('dram__bytes_write.sum', 0) / 1e6:.1f}\n")
    fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for lid, m in per.items():
        f = fam[family(names[lid])]
        f[0] += 1
        f[1] += m.get("dram__bytes_read.sum", 0.0)
        f[2] += m.get("dram__bytes_write.sum", 0.0)
        f[3] += m.get("gpu__time_duration.sum", 0.0)
    total_ms = sum(v[3] for v in fam.values())
    out = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over {len(per)} launches of `bench.py --steps 1` "
                     f"({src}); cold-cache serialised replays: shares, not absolute times, are comparable with the CUDA-event profile",
           "families": {k: {"launches": v[0], "dram_read_mb": v[1] / 1e6, "dram_write_mb": v[2] / 1e6, "ms": v[3], "share_of_time": v[3] / total_ms,
                            "avg_bytes_per_launch": (v[1] + v[2]) / v[0], "gbs": (v[1] + v[2]) / 1e6 / v[3] if v[3] else None} for k, v in fam.items()}}
    g = out["families"].get("gemm_kernel")
    if g:
        out["gemm_kernel_avg_bytes_per_launch"] = g["avg_bytes_per_launch"]
    json.dump(out, open(dst + ".json", "w"), indent=1)
    with open(dst + ".txt", "w") as f:
        f.write("# " + out["source"] + "\n")
        f.write(f"# {'kernel family':24s} {'launches':>8s} {'read MB':>10s} {'write MB':>10s} {'ms':>9s} {'share':>7s} {'GB/s':>8s}\n")
        for k, v in sorted(out["families"].items(), key=lambda kv: -kv[1]["ms"]):
            f.write(f"{k:26s} {v['launches']:8d} {v['dram_read_mb']:10.1f} {v['dram_write_mb']:10.1f} {v['ms']:9.3f} {100 * v['share_of_time']:6.1f}% {v['gbs'] or 0:8.0f}\n")
    print(open(dst + ".txt").read())


if __name__ == "__main__":
    main(*sys.argv[1:4])

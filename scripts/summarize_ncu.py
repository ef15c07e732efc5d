"""Turn ncu reports brought back from the GPU box (gpurun_out/*.ncu-rep) into small text summaries under profiles/.

    python scripts/summarize_ncu.py gpurun_out/prof_gemm_v4.ncu-rep profiles/r01_ncu_gemm.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep, out):
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full --clock-control none  ({rep}); one block per captured launch\n"]
    for n, r in enumerate(raw[2:]):
        lines.append(f"## launch {n}: {r[idx['Kernel Name']][:110]}")
        for k in KEYS:
            if k in idx:
                lines.append(f"  {k:82s} {r[idx[k]]:>16s} {units[idx[k]]}")
        lines.append("")
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-kernel-base", "function"]))))
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    seen = set()
    for bi, b in enumerate(blocks):
        key = (b["name"], len(b["data"]), sum(int(r[b["hdr"].index("# Samples")]) for r in b["data"]))
        if key in seen or not b["data"]:
            continue
        seen.add(key)
        h = b["hdr"]
        isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        sc = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        tot = sum(int(r[isamp]) for r in b["data"]) or 1
        agg = {}
        for r in b["data"]:
            for i in sc:
                agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
        lines.append(f"## source-level sampling, capture {bi // 2}: {b['name'][:60]}  ({tot} samples)")
        lines.append("  stall mix: " + ", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
        for r in sorted(b["data"], key=lambda r: -int(r[isamp]))[:12]:
            top = max(((int(r[i] or 0), h[i]) for i in sc))
            lines.append(f"  {100 * int(r[isamp]) / tot:5.1f}%  x{r[iex]:>10s}  {r[isrc].strip()[:64]:64s} {top[1][6:]}")
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("wrote", out, len(lines), "lines")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

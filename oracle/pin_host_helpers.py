"""Golden vectors for the host-side helpers of the path, generated from the reference's OWN files (test infrastructure; build container only):

    python oracle/pin_host_helpers.py        # writes tests/golden/host_helpers.json

Sources, loaded unmodified by file path: /root/reference/src/pipelines/context.py (uniform, get_total_steps) and
/root/reference/src/pipelines/utils.py (linear, slerp).  Both are numpy / torch only.  tests/test_host_helpers_cpu.py checks
humanvid_b200.pipeline against the committed file (the GPU box has no /root/reference).
"""
import importlib.util
import itertools
import json
import os

import torch

REF = "/root/reference/src/pipelines"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "host_helpers.json")


def load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ctx, utl = load("context"), load("utils")
    windows = []
    for step, nf, cs, stride, ov, closed in itertools.product((0, 1, 3, 6), (8, 16, 24, 30, 48, 72, 100), (16, 24), (1, 2, 3), (0, 4, 8), (True, False)):
        windows.append({"args": [step, 25, nf, cs, stride, ov, closed], "windows": [list(map(int, w)) for w in ctx.uniform(step, 25, nf, cs, stride, ov, closed)]})
    totals = []
    for nf, cs, stride, ov in ((48, 24, 1, 4), (72, 24, 2, 4), (24, 24, 1, 4), (100, 16, 3, 8)):
        totals.append({"args": [list(range(5)), 25, nf, cs, stride, ov], "total": int(ctx.get_total_steps(ctx.uniform, list(range(5)), 25, nf, cs, stride, ov))})
    interp = []
    g = torch.Generator().manual_seed(7)
    for case in range(6):
        v0 = torch.randn(2, 3, 4, generator=g, dtype=torch.float64)
        v1 = torch.randn(2, 3, 4, generator=g, dtype=torch.float64) if case < 4 else v0 * (1.0 + 0.01 * case) + 1e-3 * torch.randn(2, 3, 4, generator=g, dtype=torch.float64)
        for t in (0.0, 0.25, 0.5, 0.9):
            interp.append({"v0": v0.flatten().tolist(), "v1": v1.flatten().tolist(), "shape": [2, 3, 4], "t": t,
                           "linear": utl.linear(v0, v1, t).flatten().tolist(), "slerp": utl.slerp(v0, v1, t).flatten().tolist()})
    json.dump({"source": "src/pipelines/context.py, src/pipelines/utils.py of zhenzhiwang/HumanVid, unmodified", "windows": windows, "total_steps": totals,
               "interp": interp}, open(OUT, "w"))
    print(f"wrote {OUT}: {len(windows)} window cases, {len(totals)} totals, {len(interp)} interpolation cases, {os.path.getsize(OUT) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()

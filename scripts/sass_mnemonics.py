"""Blackwell-specific SASS mnemonics per kernel of the shipped library (static instruction counts) -> profiles/r0N_sass_mnemonics.txt.
Runs on the CPU container: cuobjdump -sass humanvid_b200/lib/libhv_b200.so.

    python scripts/sass_mnemonics.py profiles/r02_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "humanvid_b200", "lib", "libhv_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "UTMACMDFLUSH", "SYNCS", "ELECT", "HMMA", "LDSM",
        "MUFU.EX2", "FFMA2", "FMUL2", "FADD2", "FHADD", "FMNMX3", "LDGSTS", "REDUX"]


def main(out):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per, cur, order = collections.defaultdict(collections.Counter), None, []
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            d = re.sub(r"\(anonymous namespace\)::", "", d)
            d = re.sub(r"^void ", "", d)
            cur = re.sub(r"\(.*$", "", d).replace("hv::", "")
            if cur not in order:
                order.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1)
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    per[cur][k] += 1
    total = collections.Counter()
    lines = ["# cuobjdump -sass humanvid_b200/lib/libhv_b200.so : Blackwell-specific SASS mnemonics per kernel (counts of static instructions)",
             "# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG = TMA tensor load/store, UTMAPF = TMA L2 prefetch,",
             "# SYNCS = mbarrier, HMMA/LDSM = mma.sync/ldmatrix (temporal attention, small-channel convs), MUFU.EX2 = ex2.approx, LDGSTS = cp.async,",
             "# FFMA2/FMUL2/FADD2 = packed f32x2 math, FHADD = add.f32.f16, FMNMX3 = 3-input min/max", ""]
    for k in sorted(order):
        c = per.get(k)
        if not c:
            continue
        total.update(c)
        lines.append(f"{k[:76]:78s} " + "  ".join(f"{n}:{c[n]}" for n in KEYS if c[n]))
    lines += ["", "TOTAL  " + "  ".join(f"{n}:{total[n]}" for n in KEYS if total[n])]
    open(out, "w").write("\n".join(lines) + "\n")
    print(lines[-1])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_mnemonics.txt"))

#!/usr/bin/env python
"""cuobjdump -sass opcode histogram per kernel of libgraspldm_b200.so -> profiles/rNN_sass_histogram.md
(the tcgen05 / TMEM / TMA evidence: UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor,
UBLKCP = cp.async.bulk, UTMAPF = bulk prefetch, SYNCS = mbarrier, REDUX = warp reductions)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "graspldm_b200", "libgraspldm_b200.so")
KEY = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTMAPF", "UTCBAR", "SYNCS", "REDUX", "MUFU", "FFMA2", "FADD2", "FMUL2",
       "FFMA", "HFMA2", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "VOTE", "MATCH", "ATOM", "RED", "LDL", "STL"]


def _short(name):
    """demangled signature -> function name with its template arguments, without the parameter list"""
    name = name.replace("void ", "").replace("gldm::", "")
    depth = 0
    for i, ch in enumerate(name):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            return name[:i].replace("(int)", "").replace("(bool)", "")
    return name


def main(out):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    dem = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.split("\n")
    rows = []
    for (mangled, c), name in zip(kernels.items(), dem):
        short = _short(name)
        rows.append((short, sum(c.values()), c))
    rows.sort(key=lambda r: -r[1])
    lines = ["# SASS opcode histogram per kernel (`cuobjdump -sass graspldm_b200/libgraspldm_b200.so`, sm_100a)", "",
             "Static instruction counts.  `UTCHMMA` = tcgen05.mma, `LDTM` / `STTM` = tcgen05.ld / st (TMEM), `UTMALDG` = "
             "cp.async.bulk.tensor (TMA tensor load), `UBLKCP` = cp.async.bulk (TMA 1-D), `UTMAPF` = bulk L2 prefetch, `SYNCS` = "
             "mbarrier operations, `REDUX` = warp-wide integer reductions, `FFMA2 / FADD2 / FMUL2` = packed f32x2 arithmetic.", "",
             "| kernel | instructions | " + " | ".join(KEY) + " |", "|---|---:|" + "---:|" * len(KEY)]
    for short, tot, c in rows:
        lines.append(f"| `{short}` | {tot} | " + " | ".join(str(c.get(k, 0) or "") for k in KEY) + " |")
    tc = [r[0] for r in rows if r[2].get("UTCHMMA")]
    lines += ["", "Kernels that issue tcgen05.mma: " + ", ".join(f"`{k}`" for k in tc) + ".", ""]
    open(out, "w").write("\n".join(lines))
    print(f"{len(rows)} kernels -> {out}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_histogram.md"))

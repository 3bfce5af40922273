#!/usr/bin/env python
"""Summarise ncu output into small tracked files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/x_launches.csv profiles/rNN_launches.md
  python tools/ncu_summary.py full gpurun_out/x.ncu-rep profiles/rNN_kernel.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

FULL_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src, errors="replace")) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        u = r[mu]
        ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
        name = r[kn].split("(")[0].strip()
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
        total += ms
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` - per-launch times are "
                "cold-cache and serialised: compare SHARES, not absolutes.\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {ms:.3f} | {100 * ms / total:.1f}% |\n")
        f.write(f"| **total** | {sum(a[0] for a in agg.values())} | {total:.3f} | 100% |\n")


def full(src, dst):
    # src: a .ncu-rep, or the `ncu -i x.ncu-rep --page raw --csv` export of one (made on the GPU box when the report
    # itself is too large to bring back)
    if src.endswith(".csv"):
        out = open(src, errors="replace").read()
    else:
        out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(out)) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hi], rows[hi + 1]
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n`ncu --set full --clock-control none`: one replayed launch per kernel, "
                "cold caches; durations are not bench values.\n\n")
        for r in rows[hi + 2:]:
            f.write(f"## `{r[hdr.index('Kernel Name')]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in FULL_METRICS:
                if m in hdr:
                    f.write(f"| {m} | {r[hdr.index(m)]} | {units[hdr.index(m)]} |\n")
            top = []
            for h in stall:
                try:
                    top.append((float(r[hdr.index(h)].replace(",", "")), h))
                except ValueError:
                    pass
            top.sort(reverse=True)
            if top:
                f.write("\nwarp stall reasons (warps stalled per issue-active cycle, top 6): "
                        + ", ".join(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}" for v, h in top[:6]) + "\n")
            f.write("\n")


TABLE = [("time us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
         ("regs", "launch__registers_per_thread"), ("DRAM rd", "dram__bytes_read.sum"), ("DRAM wr", "dram__bytes_write.sum"),
         ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
         ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
         ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active")]


def table(src, dst):
    """One row per captured launch (raw-page CSV export): the compact form for multi-kernel passes."""
    rows = [r for r in csv.reader(open(src, errors="replace")) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hi], rows[hi + 1]
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full, one row per launch ({src})\n\n`ncu --set full --clock-control none`; cold caches, replayed launches: "
                "durations are not bench values.  Units as ncu printed them.\n\n| # | kernel | " + " | ".join(t for t, _ in TABLE)
                + " | top stalls |\n|---:|---|" + "---:|" * len(TABLE) + "---|\n")
        for n, r in enumerate(rows[hi + 2:]):
            cells = []
            for t, m in TABLE:
                if m not in hdr:
                    cells.append("")
                    continue
                v, u = r[hdr.index(m)], units[hdr.index(m)]
                try:
                    x = float(v.replace(",", ""))
                    if t == "time us":
                        x = x * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u.replace("second", "s").replace("nsecond", "ns"), 1.0) if u in ("ns", "us", "ms", "s") else x
                        v = f"{x:.1f}"
                    elif "bytes" in m:
                        v = f"{x:.2f} {u.replace('byte', 'B')}"
                    elif t in ("grid", "block", "regs"):
                        v = f"{int(x)}"
                    else:
                        v = f"{x:.1f}"
                except ValueError:
                    pass
                cells.append(v)
            top = sorted(((float(r[hdr.index(h)].replace(",", "")), h) for h in stall), reverse=True)[:3]
            st = ", ".join(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.1f}" for v, h in top)
            f.write(f"| {n} | `{r[hdr.index('Kernel Name')].split('(')[0]}` | " + " | ".join(cells) + f" | {st} |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full, "table": table}[sys.argv[1]](sys.argv[2], sys.argv[3])

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
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for r in rows[2:]:
            f.write(f"## `{r[hdr.index('Kernel Name')]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in FULL_METRICS:
                if m in hdr:
                    f.write(f"| {m} | {r[hdr.index(m)]} | {units[hdr.index(m)]} |\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])

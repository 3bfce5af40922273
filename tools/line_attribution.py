"""Development helper: join an `ncu --page source --csv` export (SASS view, per-instruction counters) with
`nvdisasm --print-line-info` of the same build, and report executed warp instructions / stall samples per source line
and per source function region.

  python tools/line_attribution.py <source.csv> <disasm.sass> <mangled kernel name> [first line-last line=label ...]
"""
import csv, re, sys, collections

src_csv, sass, kern = sys.argv[1:4]
regions = []
for a in sys.argv[4:]:
    rng, label = a.split("=")
    lo, hi = rng.split("-")
    regions.append((int(lo), int(hi), label))
cur, lines, on = None, [], False
for ln in open(sass):
    if ln.startswith(".text."):
        on = ln.strip() == f".text.{kern}:"
        continue
    if not on:
        continue
    if ln.startswith("//-----"):
        on = False
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines.append((cur, m.group(2)))
rows = list(csv.reader(open(src_csv)))
hdr, body = rows[1], rows[2:]
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
assert len(body) == len(lines), (len(body), len(lines))
byline, samp = collections.Counter(), collections.Counter()
tot = tots = 0
for (loc, txt), r in zip(lines, body):
    n, s = int(r[iI]), int(r[iS])
    byline[loc] += n; samp[loc] += s; tot += n; tots += s
print(f"{len(lines)} SASS instructions, {tot} executed warp instructions, {tots} stall samples")
if regions:
    agg = collections.Counter(); aggs = collections.Counter()
    for loc, n in byline.items():
        lab = "other"
        if loc and loc[0] == "sampler_rows.cuh":
            for lo, hi, label in regions:
                if lo <= loc[1] <= hi:
                    lab = label
                    break
        elif loc:
            lab = loc[0]
        agg[lab] += n; aggs[lab] += samp[loc]
    for lab, n in agg.most_common():
        print(f"{100 * n / tot:5.1f}% instr  {100 * aggs[lab] / max(tots, 1):5.1f}% samples  {lab}")
print("top lines:")
for loc, n in byline.most_common(40):
    print(f"{100 * n / tot:5.1f}%  {n:9d}  samples {100 * samp[loc] / max(tots, 1):4.1f}%  {loc}")

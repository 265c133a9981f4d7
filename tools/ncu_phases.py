#!/usr/bin/env python
"""Phase breakdown of k_deposit_tile5 from an ncu source dump (cuda,sass): per SASS instruction, the CUDA line of
deposit_tile5.cuh it was attributed to, bucketed by line ranges.  Inlined helpers are attributed to the bucket of the
nearest preceding tile5 line in address order."""
import csv, sys
PH = [(0, 179, "headers+setup"), (180, 225, "extent+prefetch"), (226, 359, "classification"), (360, 460, "run table + cov/dels"),
      (461, 526, "window/slab setup"), (527, 645, "staging (keys)"), (646, 677, "task fetch"), (678, 710, "pass loop"),
      (711, 767, "bit-sliced reduce"), (768, 834, "flush + first-seen"), (835, 9999, "deferred/tail")]
def phase(ln):
    for a, b, n in PH:
        if a <= ln <= b: return n
    return "?"
rows = []
cur_file = None; line_no = None; ix = None
seen = set()
for row in csv.reader(open(sys.argv[1], newline="")):
    if not row: continue
    if row[0] == "File Path": cur_file = row[1].split("/")[-1]; continue
    if row[0] == "Line No": ix = {n: i for i, n in enumerate(row)}; continue
    if row[0] == "Function Name" or ix is None: continue
    if row[0] != "": line_no = int(row[0]); continue
    if row[2] == "..." : continue
    addr = row[2]
    try: inst = int(float(row[ix["Instructions Executed"]])); smp = int(float(row[ix["# Samples"]]))
    except ValueError: continue
    rows.append((int(addr, 16), cur_file, line_no, inst, smp, row[3]))
# an address appears once per inlining level: keep the tile5 attribution if any, else the first one
by_addr = {}
for a, f, ln, inst, smp, sass in rows:
    if a not in by_addr or (f == "deposit_tile5.cuh" and by_addr[a][0] != "deposit_tile5.cuh"):
        by_addr[a] = (f, ln, inst, smp, sass)
tot_i = sum(v[2] for v in by_addr.values()); tot_s = sum(v[3] for v in by_addr.values())
agg = {}
last = "headers+setup"
for a in sorted(by_addr):
    f, ln, inst, smp, sass = by_addr[a]
    if f == "deposit_tile5.cuh": last = phase(ln)
    p = last
    x = agg.setdefault(p, [0, 0]); x[0] += inst; x[1] += smp
print(f"warp instructions {tot_i}, samples {tot_s}")
for a, b, n in PH:
    if n in agg: print(f"{n:28s} {100*agg[n][0]/tot_i:5.1f}% inst  {100*agg[n][1]/tot_s:5.1f}% samples")

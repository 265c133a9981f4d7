#!/usr/bin/env python
"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` dump per CUDA source line:
warp instructions executed, stall samples, shared-memory wavefronts.  Usage: ncu_lines.py dump.csv [top_n]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cur_file, hdr = None, None
    agg = defaultdict(lambda: [0, 0, 0, 0, ""])      # (file, line) -> inst, samples, smem wavefronts, excessive, text
    line_no, line_src = None, ""
    for row in csv.reader(open(path, newline="")):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file = row[1].split("/")[-1]
            continue
        if row[0] == "Line No":
            hdr = row
            ix = {n: i for i, n in enumerate(hdr)}
            continue
        if row[0] == "Function Name" or hdr is None:
            continue
        if row[0] != "":
            line_no, line_src = row[0], row[1]
            continue
        if row[2] == "...":
            continue

        def num(name):
            try:
                return int(float(row[ix[name]]))
            except (ValueError, KeyError):
                return 0
        a = agg[(cur_file, int(line_no))]
        a[0] += num("Instructions Executed")
        a[1] += num("# Samples")
        a[2] += num("L1 Wavefronts Shared")
        a[3] += num("L1 Wavefronts Shared Excessive")
        a[4] = line_src.strip()[:110]
    tot_i = sum(a[0] for a in agg.values()) or 1
    tot_s = sum(a[1] for a in agg.values()) or 1
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    print("--- by instructions")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*a[0]/tot_i:5.1f}% inst {100*a[1]/tot_s:5.1f}% smp  wf {a[2]:>9} exc {a[3]:>8}  {f}:{ln}  {a[4]}")
    print("--- by stall samples")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100*a[1]/tot_s:5.1f}% smp {100*a[0]/tot_i:5.1f}% inst  {f}:{ln}  {a[4]}")


if __name__ == "__main__":
    main()

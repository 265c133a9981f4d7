#!/usr/bin/env python
"""SASS opcode histogram per kernel of the built library (cuobjdump -sass), as markdown.
Usage: sass_hist.py [lib.so] > profiles/r2_sass_histogram.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "covid-spings-variant-caller_b200", "lvc_b200", "liblvc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist, order = None, {}, []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter(); order.append(kern)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
print("# SASS opcode histogram per kernel (`cuobjdump -sass liblvc_b200.so`, sm_100a cubin)\n")
print("Base mnemonics (modifiers folded).  `UBLKCP` + `SYNCS` = TMA bulk copy + mbarrier (the TMA staging variant `k_deposit_tile`);\n`ATOMS`/`ATOMG`/`RED` = shared / global atomics and reductions; `DFMA`/`DADD`/`DMUL` = fp64 (genotype only).\n")
for k in order:
    c = hist[k]
    tot = sum(c.values())
    print(f"## `{k}` — {tot} instructions\n")
    print(", ".join(f"{op} {n}" for op, n in c.most_common()))
    print()

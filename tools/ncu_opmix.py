#!/usr/bin/env python3
"""Executed-instruction mix by full SASS opcode from an ncu report.  usage: ncu_opmix.py <report> [top]"""
import collections, csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; isrc, ii = hdr.index("Source"), hdr.index("Instructions Executed")
c = collections.Counter(); tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    toks = r[isrc].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    n = int(float(r[ii] or 0)); c[op] += n; tot += n
print("total", f"{tot:.3e}")
for op, n in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print(f"{op:28s} {100*n/tot:5.1f}%")

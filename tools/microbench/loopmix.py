#!/usr/bin/env python3
"""Count SASS opcodes inside the innermost backward-branch loop of every kernel in a cubin/exe."""
import re, subprocess, sys, collections
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
fn = None; ins = []
def flush():
    if fn is None or not ins: return
    # find last backward BRA: target address < own address
    loops = []
    for idx, (addr, op, rest) in enumerate(ins):
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m and int(m.group(1), 16) < addr: loops.append((int(m.group(1), 16), addr))
    if not loops: return
    lo, hi = max(loops, key=lambda t: t[1] - t[0])
    c = collections.Counter(op for a, op, r in ins if lo <= a <= hi)
    print(fn, dict(sorted(c.items(), key=lambda kv: -kv[1])))
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush(); fn = m.group(1); ins = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)\s*(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
flush()

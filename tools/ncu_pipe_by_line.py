#!/usr/bin/env python3
"""Per source line: executed warp-instructions on the IMAD (fmaheavy) pipe vs other.  usage: <report> <lib.so> <kernel> [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
addr2line, cur, insec = {}, None, False
for ln in dis.splitlines():
    if ln.startswith("//--------------------- .text."):
        insec = kern in ln; continue
    if not insec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m: addr2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ia, isrc, ii = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
agg = collections.defaultdict(lambda: [0.0, 0]); base = None; tw = tn = 0
W = {"IMAD.HI": 2.19, "IMAD.WIDE": 2.77}
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None: base = a
    toks = r[isrc].split(); op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    n = int(float(r[ii] or 0)); key = addr2line.get(a - base)
    agg[key][1] += n; tn += n
    if op.startswith("IMAD") or op.startswith("IMUL"):
        w = next((v for k, v in W.items() if op.startswith(k)), 1.0)
        agg[key][0] += w * n; tw += w * n
print(f"IMAD-pipe weighted slots {tw:.3e} of {tn:.3e} warp-instructions")
for key, g in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(key):38s} imad_slots {100*g[0]/tw:5.1f}%   inst {100*g[1]/tn:5.1f}%")

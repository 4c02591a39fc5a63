#!/usr/bin/env python3
"""Join an ncu SASS-level source page with nvdisasm line info: per source line, instructions executed and
stall samples.   usage: ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-substring> [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
addr2line, cur, insec = {}, None, False
for ln in dis.splitlines():
    if ln.startswith("//--------------------- .text."):
        insec = kern in ln
        continue
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iconf = hdr.index("L1 Wavefronts Shared Excessive")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter(), collections.Counter()])
tot_i = tot_s = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = a
    key, ins = addr2line.get(a - base, (None, "?"))
    n, s = int(float(r[ii] or 0)), int(float(r[isamp] or 0))
    g = agg[key]
    g[0] += n; g[1] += s; g[2] += int(float(r[iconf] or 0))
    g[3][ins.split()[0].split(".")[0] if ins != "?" else "?"] += n
    for i, h in stall_cols:
        v = int(float(r[i] or 0))
        if v:
            g[4][h] += v
    tot_i += n; tot_s += s
print(f"total warp-instructions {tot_i:.3e}  samples {tot_s}")
for key, g in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    ops = " ".join(f"{k}:{100*v/max(g[0],1):.0f}%" for k, v in g[3].most_common(4))
    st = " ".join(f"{k[6:]}:{100*v/max(g[1],1):.0f}%" for k, v in g[4].most_common(3))
    print(f"{str(key):38s} inst {100*g[0]/tot_i:5.1f}%  samples {100*g[1]/tot_s:5.1f}%  smem_excess {g[2]:.2e} | {ops} | {st}")

#!/usr/bin/env python3
"""Shared-memory wavefronts per source line: ideal vs excessive (bank conflicts), from an ncu source page.
usage: smem_by_line.py <report.ncu-rep> <cubin> <kernel-substring>"""
import csv, re, subprocess, sys
from collections import defaultdict
rep, cubin, kname = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
start = next((i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kname in ",".join(r)), 0)
hdr_i = next(i for i, r in enumerate(rows) if i >= start and r and r[0] == "Address")
hdr = rows[hdr_i]
iw, ii, ie, ix = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal"), hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("Instructions Executed")
insts = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    insts.append((r[1].strip(), int(r[iw] or 0), int(r[ii] or 0), int(r[ie] or 0), int(r[ix] or 0)))
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
lines, cur, infunc = [], None, False
for l in dis.splitlines():
    if l.startswith("//--------------------- .text."):
        infunc = kname in l
        continue
    if not infunc:
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        lines.append(cur)
per = defaultdict(lambda: [0, 0, 0, 0, set()])
for (txt, w, i, e, x), ln in zip(insts, lines):
    if w:
        p = per[ln]
        p[0] += w; p[1] += i; p[2] += e; p[3] += x; p[4].add(txt.split()[0] if not txt.startswith('@') else txt.split()[1])
tw = sum(p[0] for p in per.values())
print(f"shared wavefronts total {tw}")
for ln, (w, i, e, x, ops) in sorted(per.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{str(ln):34s} wavefronts {w:10d} ({100*w/tw:5.1f}%) ideal {i:10d} excess {e:10d}  warp-instr {x:9d}  x{w/max(x,1):.1f}  {','.join(sorted(ops))}")

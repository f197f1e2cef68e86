#!/usr/bin/env python3
"""Join an ncu SASS-level source page (per-instruction executed counts) with nvdisasm line info,
and print executed warp-instructions per CUDA source line / per source-line range.

usage: sass_by_line.py <report.ncu-rep> <cubin> <kernel-substring> [top_n]
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# the block of the kernel whose name contains <kernel-substring> (first match)
start = next((i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kname in ",".join(r)), 0)
hdr_i = next(i for i, r in enumerate(rows) if i >= start and r and r[0] == "Address")
hdr = rows[hdr_i]
ie = hdr.index("Instructions Executed")
ist = hdr.index("Warp Stall Sampling (All Samples)")
insts = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    if len(r) > ie:
        insts.append((r[1].strip(), int(r[ie] or 0), int(r[ist] or 0)))
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
print(f"sass instructions: report {len(insts)}, disasm {len(lines)}")
n = min(len(insts), len(lines))
per = defaultdict(lambda: [0, 0, 0])
tot = sum(c for _, c, _ in insts)
tots = sum(s for _, _, s in insts)
for (txt, cnt, st), ln in zip(insts[:n], lines[:n]):
    per[ln][0] += cnt
    per[ln][1] += st
    per[ln][2] += 1
print(f"total warp instructions executed: {tot}, stall samples {tots}")
for ln, (c, s, k) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(ln):34s} exec {c:12d} ({100*c/tot:5.1f}%)  stalls {100*s/max(tots,1):5.1f}%  sass {k}")
if '--by-stall' in sys.argv:
    print('--- by stall samples')
    for ln, (c, s, k) in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{str(ln):34s} exec {c:12d} ({100*c/tot:5.1f}%)  stalls {100*s/max(tots,1):5.1f}%  sass {k}")

if len(sys.argv) > 5:
    # ranges "name:lo-hi,name:lo-hi" over the main .cu file; helper headers are attributed by name
    groups = defaultdict(int)
    spec = [(g.split(":")[0], *map(int, g.split(":")[1].split("-"))) for g in sys.argv[5].split(",")]
    for ln, (c, s, k) in per.items():
        name = "other"
        if ln is None:
            name = "noline"
        elif ln[0].endswith(".cu"):
            for nm, lo, hi in spec:
                if lo <= ln[1] <= hi:
                    name = nm
                    break
        else:
            name = ln[0]
        groups[name] += c
    for nm, c in sorted(groups.items(), key=lambda kv: -kv[1]):
        print(f"{nm:28s} {c:12d} {100*c/tot:5.1f}%")

#!/usr/bin/env python
"""Join an `ncu --page source --csv --print-source sass` dump with `nvdisasm -g` line info of the same cubin and
aggregate stall samples / executed instructions per CUDA source line.
usage: python tools/sass_lines.py X_sass.csv <cubin> <mangled kernel name substring> [topN]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
dis = subprocess.run(["nvdisasm", "-g", sys.argv[2]], capture_output=True, text=True).stdout.splitlines()
key = sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
lines, cur, on = [], None, False
for l in dis:
    if l.startswith("//---") and ".text." in l:
        on = key in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
assert len(lines) == len(body), (len(lines), len(body))
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
tot = 0
for ln, r in zip(lines, body):
    s, n = int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0)
    a = agg[ln]
    a[0] += s
    a[1] += n
    tot += s
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            a[2][k[6:]] += int(r[ix[k]] or 0)
src = {}
for ln, (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, no = ln if ln else ("?", 0)
    if f not in src:
        try:
            src[f] = open(f"/root/repo/cddmsl_b200/csrc/{f}").read().splitlines()
        except OSError:
            src[f] = []
    text = src[f][no - 1].strip()[:70] if 0 < no <= len(src[f]) else ""
    tops = ", ".join(f"{k}={100*v/max(s,1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*s/max(tot,1):5.1f}%  inst={n:>12,}  {f}:{no:<4d} {text}   [{tops}]")

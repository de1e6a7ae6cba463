#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` dump: instruction mix, stall samples by opcode, and
the hottest instructions.  usage: python tools/ncu_sass_top.py X_sass.csv [topN]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ins, samples, stall = Counter(), Counter(), {}
recs = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0] + ("." + src.split(".")[1].split()[0] if op in ("LDG", "STG", "LDS", "STS", "BRA") and "." in src.split()[0 if not src.startswith("@") else 1] else "")
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    ins[op] += n
    samples[op] += s
    recs.append((s, n, r[ix["Address"]], src, {k: int(r[ix[k]] or 0) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}))
tot_i, tot_s = sum(ins.values()), sum(samples.values())
print(f"instructions {tot_i:,}  samples {tot_s:,}  static {len(recs)}")
print("opcode            inst%  samples%")
for op, n in ins.most_common(22):
    print(f"  {op:14s} {100*n/tot_i:6.2f} {100*samples[op]/max(tot_s,1):8.2f}")
agg = Counter()
for s, n, a, src, st in recs:
    for k, v in st.items():
        agg[k] += v
print("stall totals:", ", ".join(f"{k[6:]}={100*v/max(tot_s,1):.1f}%" for k, v in agg.most_common(8)))
print("hottest instructions:")
for s, n, a, src, st in sorted(recs, key=lambda t: -t[0])[:top]:
    main = max(st.items(), key=lambda kv: kv[1])
    print(f"  {100*s/max(tot_s,1):5.2f}%  n={n:>11,}  {main[0][6:]:10s} {src[:90]}")

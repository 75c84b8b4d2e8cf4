#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: stall totals and the hottest SASS lines."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = {s: 0 for s in stalls}
lines = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[0] == "Address": continue
    samples = int(r[col["# Samples"]] or 0)
    for s in stalls: tot[s] += int(r[col[s]] or 0)
    lines.append((samples, r[col["Source"]].strip(), int(r[col["Instructions Executed"]] or 0),
                  {s: int(r[col[s]] or 0) for s in stalls}))
all_s = sum(tot.values())
print("total samples", all_s, " SASS lines", len(lines), " warp-instr", sum(l[2] for l in lines))
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {s:24s} {v:8d} {100.0*v/max(all_s,1):5.1f}%")
print("hottest lines:")
for samples, src, n, st in sorted(lines, key=lambda l: -l[0])[:top]:
    main = max(st.items(), key=lambda kv: kv[1])
    print(f"  {samples:7d} {100.0*samples/max(all_s,1):5.1f}% exec={n:9d} {main[0]:18s} {src[:90]}")

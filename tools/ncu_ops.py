#!/usr/bin/env python
"""Dynamic per-opcode instruction counts from `ncu -i rep --page source --csv`, per work item.
Usage: ncu_ops.py report.ncu-rep n_items [top]"""
import collections, csv, io, subprocess, sys
rep, n = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; col = {k: i for i, k in enumerate(hdr)}
ops = collections.Counter(); tot = 0
for r in rows[hi + 1:]:
    if len(r) != len(hdr): continue
    k = int(r[col["Instructions Executed"]] or 0)
    toks = r[col["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    ops[op.split(".")[0]] += k; tot += k
print(f"warp instructions {tot}  = {tot*32/n:.1f} thread-instructions per item")
for k, v in ops.most_common(top): print(f"  {k:10s} {v:11d} {v*32/n:7.2f}/item")

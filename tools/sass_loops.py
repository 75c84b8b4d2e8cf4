#!/usr/bin/env python
"""Static SASS loop report: for each backward branch of a cuobjdump -sass dump, the opcode
histogram of the loop body (largest bodies first).  Usage: sass_loops.py dump.txt [min_len]"""
import collections, re, sys
lines = []
for l in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((int(m.group(1), 16), m.group(2).strip()))
min_len = int(sys.argv[2]) if len(sys.argv) > 2 else 40
addr_index = {a: i for i, (a, _) in enumerate(lines)}
loops = []
for i, (a, ins) in enumerate(lines):
    m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?\d*\)?\s*$", ins)
    m2 = re.search(r"BRA.*0x([0-9a-f]+)", ins)
    if m2:
        t = int(m2.group(1), 16)
        if t <= a and t in addr_index:
            loops.append((addr_index[t], i))
for s, e in sorted(loops, key=lambda x: x[0] - x[1]):
    if e - s < min_len:
        continue
    ops = collections.Counter()
    for a, ins in lines[s:e + 1]:
        toks = ins.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ops[op.split(".")[0]] += 1
    print(f"loop {lines[s][0]:#x}..{lines[e][0]:#x}: {e - s + 1} instructions")
    print("   " + "  ".join(f"{k}:{v}" for k, v in ops.most_common()))

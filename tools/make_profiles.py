#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call left in gpurun_out/ into the tracked summaries under
profiles/.   Usage: make_profiles.py <tag in gpurun_out, e.g. r1e> <round label, e.g. r01>
Reads   gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_{nn,nn_tc,picp,picpres,tri}.ncu-rep
Writes  profiles/<label>_launches.csv, profiles/<label>_launches.md, profiles/<label>_<k>.md"""
import collections, csv, io, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, label = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def run(*a):
    return subprocess.run(a, capture_output=True, text=True).stdout


lc = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(lc):
    shutil.copy(lc, os.path.join(P, f"{label}_launches.csv"))
    rows = list(csv.reader(open(lc)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]; col = {k: i for i, k in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) != len(hdr): continue
        name = r[col["Kernel Name"]].split("(")[0]
        a = agg.setdefault(name[-70:], [0, 0.0]); a[0] += 1; a[1] += float(r[col["Metric Value"]])
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f"{label}_launches.md"), "w") as f:
        f.write(f"# {label}: launch list of one `bench.py` run under `ncu --metrics gpu__time_duration.sum "
                "--clock-control none` (cold-cache, serialised: shares, not absolutes)\n\n"
                "| launches | total ms | share | kernel |\n|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
            f.write(f"| {n} | {t/1e6:.3f} | {100*t/tot:.1f} % | `{k}` |\n")

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
traffic = {}
for k in ("nn", "nn_tc", "picp", "picpres", "tri"):
    rep = os.path.join(G, f"{tag}_{k}.ncu-rep")
    if not os.path.exists(rep): continue
    raw = list(csv.reader(io.StringIO(run("ncu", "-i", rep, "--page", "raw", "--csv"))))
    col = dict(zip(raw[0], zip(raw[1], raw[2])))
    rd = float(col["dram__bytes_read.sum"][1]) * UNIT[col["dram__bytes_read.sum"][0]]
    wr = float(col["dram__bytes_write.sum"][1]) * UNIT[col["dram__bytes_write.sum"][0]]
    traffic[k] = {"kernel": col["Kernel Name"][1], "grid": col["Grid Size"][1],
                  "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
                  "gpu_time_ns_under_ncu": float(col["gpu__time_duration.sum"][1]) *
                  {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}[col["gpu__time_duration.sum"][0]],
                  "report": f"gpurun_out/{tag}_{k}.ncu-rep"}
for k in ("nn", "nn_tc", "picp", "picpres", "tri"):
    rep = os.path.join(G, f"{tag}_{k}.ncu-rep")
    if not os.path.exists(rep): continue
    out = [f"# {label}: `ncu --set full --clock-control none --import-source on` of the {k} kernel\n",
           f"(source report: gpurun_out/{tag}_{k}.ncu-rep, summarised by tools/make_profiles.py)\n",
           run(sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep)]
    src = run("ncu", "-i", rep, "--page", "source", "--csv")
    tmp = f"/tmp/{tag}_{k}_src.csv"
    open(tmp, "w").write(src)
    out.append("## warp stall sampling and hottest SASS lines\n\n```\n" +
               run(sys.executable, os.path.join(ROOT, "tools", "ncu_src.py"), tmp, "12") + "```\n")
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]; col = {n: i for i, n in enumerate(hdr)}
    ops = collections.Counter(); tot = 0
    for r in rows[hi + 1:]:
        if len(r) != len(hdr): continue
        n = int(r[col["Instructions Executed"]] or 0)
        toks = r[col["Source"]].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ops[op.split(".")[0]] += n; tot += n
    out.append("## executed warp-instructions by opcode\n\n| opcode | warp-instructions | share |\n|---|---|---|\n" +
               "".join(f"| {o} | {v} | {100*v/tot:.1f} % |\n" for o, v in ops.most_common(14)))
    open(os.path.join(P, f"{label}_ncu_{k}.md"), "w").write("\n".join(out))
if traffic:
    import json
    json.dump(traffic, open(os.path.join(P, f"{label}_traffic.json"), "w"), indent=1)
print(sorted(os.listdir(P)))

#!/usr/bin/env python
"""Randomised agreement check of the two exact NN paths (tensor-core filter vs FP32 filter, both
forced) over random map sizes, batch sizes, radii, coordinate scales and match densities; every
disagreement in index or d2 is printed.   python tools/nn_stress.py [n_cases] [seed]   (needs a GPU)"""
import importlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    vo = importlib.import_module("visual-odometry_b200")
    n_cases, seed = int(sys.argv[2]), int(sys.argv[3])
    rng = np.random.RandomState(seed)
    out = []
    for case in range(n_cases):
        M = int(rng.choice([2048, 2049, 4097, 10000, 32767, 32768, 33001, 65537, 100003, 262144 + 77]))
        Q = int(rng.choice([512, 513, 1000, 2047, 2048, 2050, 4096, 9001, 20000, 33333]))
        scale = float(rng.choice([0.05, 0.5, 1.0, 3.0]))
        norm = float(rng.choice([0.05, 0.1, 0.3])) * scale
        m = (scale * rng.uniform(-1, 1, (M, 11))).astype(np.float32)
        q = (scale * rng.uniform(-1, 1, (Q, 11))).astype(np.float32)
        kind = rng.randint(0, 4, Q)                       # 0 exact copy, 1 near, 2 at the edge, 3 none
        rows = rng.randint(0, M, Q)
        d = rng.normal(size=(Q, 10)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        frac = np.choose(kind, [0.0, 0.3, 1.0 - 2e-6, 0.0])
        planted = (m[rows, 1:].astype(np.float64) + d * (norm * frac)[:, None]).astype(np.float32)
        q[kind < 3, 1:] = planted[kind < 3]
        if rng.rand() < 0.5:                              # duplicate rows: ties
            dup = rng.randint(0, M, 50)
            m[rng.randint(0, M, 50)] = m[dup]
        nn = vo.NNIndex(0)
        nn.set_map(m)
        idx, d2 = nn.best_match(q, norm, want_d2=True)
        out.append({"M": M, "Q": Q, "scale": scale, "norm": norm, "launch": nn.last_launches()[0][:3],
                    "idx": idx.tolist(), "d2": d2.view(np.uint32).tolist()})
        nn.close()
    json.dump(out, sys.stdout)
else:
    n_cases = sys.argv[1] if len(sys.argv) > 1 else "24"
    seed = sys.argv[2] if len(sys.argv) > 2 else "1"
    res = {}
    for path in ("ffma", "tc"):
        o = subprocess.run([sys.executable, __file__, "child", n_cases, seed], env=dict(os.environ, VO_NN_FORCE_PATH=path),
                           capture_output=True, text=True)
        if o.returncode:
            print(path, "FAILED", o.stderr[-500:]); sys.exit(1)
        res[path] = json.loads(o.stdout)
    bad = 0
    for a, b in zip(res["ffma"], res["tc"]):
        same_idx = a["idx"] == b["idx"]
        hit = [i for i, v in enumerate(a["idx"]) if v >= 0]
        same_d2 = all(a["d2"][i] == b["d2"][i] for i in hit)
        ok = same_idx and same_d2
        bad += 0 if ok else 1
        print(f"M={a['M']:7d} Q={a['Q']:6d} scale={a['scale']:4} norm={a['norm']:.4f} matches={len(hit):6d} "
              f"ffma={a['launch']} tc={b['launch']}  {'ok' if ok else 'MISMATCH'}")
    print("cases", len(res["ffma"]), "mismatches", bad)
    sys.exit(1 if bad else 0)

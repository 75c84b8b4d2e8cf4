"""Latency of one vo_nn_best_match call with a single query against a frame-sized map (the
reference main's access pattern, vo_complete.cpp:37-38).   python tools/nn_tiny_time.py [rows]"""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
vo = importlib.import_module("visual-odometry_b200")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.RandomState(1)
m = rng.uniform(-1, 1, (rows, 11)).astype(np.float32)
q = rng.uniform(-1, 1, (1, 11)).astype(np.float32)
nn = vo.NNIndex(0)
nn.set_map(m)
for _ in range(200):
    nn.best_match(q, 0.5)
t0 = time.perf_counter()
n = 5000
for _ in range(n):
    nn.best_match(q, 0.5)
dt = time.perf_counter() - t0
print(json.dumps({"rows": rows, "us_per_call_incl_ctypes": round(dt / n * 1e6, 2), "launches": nn.last_launches()}))

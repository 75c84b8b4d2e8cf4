#!/usr/bin/env python
"""Where does the tensor-core filter overtake the FP32 one?  set_map_device + best_match_device,
device-resident, both paths forced, over a grid of map and batch sizes (needs a GPU)."""
import importlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import numpy as np, torch
    vo = importlib.import_module("visual-odometry_b200")
    synth = importlib.import_module("visual-odometry_b200.synth")
    dev = torch.device("cuda:0")
    out = {}
    for M in (2048, 4096, 8192, 12000, 16384, 24000, 32768):
        m = torch.from_numpy(synth.nn_map_rows_np(0, M)).to(dev)
        for Q in (512, 2048, 4096, 9000, 30000):
            qn, _ = synth.nn_queries_np(Q, M)
            q = torch.from_numpy(qn).to(dev)
            idx = torch.empty(Q, dtype=torch.int32, device=dev)
            nn = vo.NNIndex(0)
            nn.set_stream(torch.cuda.current_stream().cuda_stream)
            def step():
                nn.set_map_device(m.data_ptr(), M, 11, 1)
                nn.best_match_device(q.data_ptr(), Q, 11, 0.1, idx.data_ptr())
            for _ in range(5): step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(40): step()
            e1.record(); torch.cuda.synchronize()
            out[f"{M}x{Q}"] = round(e0.elapsed_time(e1) / 40 * 1e3, 1)
            nn.close()
    print(json.dumps(out))
else:
    res = {}
    for path in ("ffma", "tc"):
        o = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, VO_NN_FORCE_PATH=path),
                           capture_output=True, text=True).stdout.strip().splitlines()[-1]
        res[path] = json.loads(o)
    print("rows x queries: us ffma / us tc")
    for k in res["ffma"]:
        print(f"{k:>14s}: {res['ffma'][k]:8.1f} / {res['tc'][k]:8.1f}  {'TC' if res['tc'][k] < res['ffma'][k] else 'ffma'}")

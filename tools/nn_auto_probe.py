import importlib, json, os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
vo = importlib.import_module("visual-odometry_b200")
synth = importlib.import_module("visual-odometry_b200.synth")
dev = torch.device("cuda:0")
for M, Q in ((4096, 4096), (8192, 9000), (10000, 10000), (12000, 4096), (16384, 9000), (30000, 30000), (32768, 2048)):
    m = torch.from_numpy(synth.nn_map_rows_np(0, M)).to(dev)
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
    print(M, Q, round(e0.elapsed_time(e1) / 40 * 1e3, 1), "us", nn.last_launches())
    nn.close()

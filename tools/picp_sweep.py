#!/usr/bin/env python
"""Config 3 (picp_test scale sweep): device-resident time per Gauss-Newton round vs N_corr.
Usage: python tools/picp_sweep.py [rounds]   (needs a GPU)"""
import importlib, json, os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vo = importlib.import_module("visual-odometry_b200")
synth = importlib.import_module("visual-odometry_b200.synth")
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
for n_gen in (1_000, 10_000, 100_000, 1_000_000, 3_000_000, 10_000_000):
    pr = synth.picp_problem(n_gen, seed=42)
    n = len(pr["pairs"])
    world = torch.from_numpy(pr["world"]).to(dev)
    image = torch.from_numpy(pr["image"]).to(dev)
    pairs = torch.from_numpy(pr["pairs"]).to(dev)
    cam = vo.Camera(pr["rows"], pr["cols"], pr["z_near"], pr["z_far"], pr["K"], np.eye(4))
    s = vo.PICPSolver(0)
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    s.setKernelThreshold(10000.0)

    def step():
        s.init_device(cam, world.data_ptr(), world.shape[0], image.data_ptr(), image.shape[0])
        s.set_correspondences_device(pairs.data_ptr(), n)
        s.compute(False, rounds)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps / rounds
    print(json.dumps({"n_corr": n, "us_per_round": round(us, 3), "GBps": round(28.0 * n / us / 1e3, 1),
                      "point_iters_per_s": n / us * 1e6}), flush=True)
    s.close()

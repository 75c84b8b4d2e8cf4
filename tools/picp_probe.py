#!/usr/bin/env python
"""One PICP compute() of `rounds` rounds on n_gen generated points (for ncu captures)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vo = importlib.import_module("visual-odometry_b200")
synth = importlib.import_module("visual-odometry_b200.synth")
n_gen, rounds = int(sys.argv[1]), int(sys.argv[2])
pr = synth.picp_problem(n_gen, seed=42)
cam = vo.Camera(pr["rows"], pr["cols"], pr["z_near"], pr["z_far"], pr["K"], np.eye(4))
s = vo.PICPSolver(0)
s.setKernelThreshold(10000.0)
s.init(cam, pr["world"], pr["image"])
s.set_correspondences(pr["pairs"])
for _ in range(3):
    s.compute(False, rounds)
print(len(pr["pairs"]), s.state().rounds_done, s.pose()[:3, 3])
s.close()

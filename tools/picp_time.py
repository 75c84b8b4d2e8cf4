"""Device-resident time per Gauss-Newton round of the streaming PICP kernel at 1e7 generated points.
   VO_B200_LIB=<variant.so> python tools/picp_time.py [n_gen] [rounds]"""
import importlib, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
vo = importlib.import_module("visual-odometry_b200")
synth = importlib.import_module("visual-odometry_b200.synth")
n_gen = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
pr = synth.picp_problem(n_gen, seed=42)
n = len(pr["pairs"])
world, image, pairs = (torch.from_numpy(pr[k]).to(dev) for k in ("world", "image", "pairs"))
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
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 10 / rounds
print(json.dumps({"lib": os.environ.get("VO_B200_LIB", "default"), "n_corr": n, "us_per_round": round(us, 3),
                  "GBps": round(28.0 * n / us / 1e3, 1), "pose_err": float(np.max(np.abs(s.pose() - pr["T_gt"])))}))

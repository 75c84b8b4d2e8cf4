"""Device-resident time of triangulate_kernel at 1e7 generated correspondences.
   VO_B200_LIB=<variant.so> python tools/tri_time.py [n]"""
import ctypes as C, importlib, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
vo = importlib.import_module("visual-odometry_b200")
synth = importlib.import_module("visual-odometry_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
tv = synth.two_view_problem(n, seed=7, noise=0.2)
nc = len(tv["corr"])
dev = torch.device("cuda:0")
corr, p1, p2 = (torch.from_numpy(tv[k]).to(dev) for k in ("corr", "p1", "p2"))
out_pts = torch.empty((nc, 3), dtype=torch.float32, device=dev)
out_cn = torch.empty((nc, 2), dtype=torch.int32, device=dev)
nsucc = torch.zeros(1, dtype=torch.int64, device=dev)
lib = vo.lib()
ws = torch.empty(int(lib.vo_triangulate_workspace_bytes(nc)), dtype=torch.uint8, device=dev)
K = np.ascontiguousarray(tv["K"].T).reshape(-1).astype(np.float32)
X = np.ascontiguousarray(tv["X"].T).reshape(-1).astype(np.float32)
f32p = C.POINTER(C.c_float)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def step():
    rc = lib.vo_triangulate_device(stream, K.ctypes.data_as(f32p), X.ctypes.data_as(f32p), C.c_void_p(corr.data_ptr()), nc,
                                   C.c_void_p(p1.data_ptr()), C.c_void_p(p2.data_ptr()), None, C.c_void_p(out_pts.data_ptr()),
                                   C.c_void_p(out_cn.data_ptr()), None, None, C.c_void_p(nsucc.data_ptr()), C.c_void_p(ws.data_ptr()))
    assert rc == 0, lib.vo_last_error()
for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(json.dumps({"lib": os.environ.get("VO_B200_LIB", "default"), "n_corr": nc, "n_success": int(nsucc.item()),
                  "ms": round(ms, 5), "GBps": round(44.0 * nc / ms / 1e6, 1)}))

"""Times one best_match_device call per filter path at a given size (device-resident inputs).
   python tools/nn_tc_time.py [M] [Q]"""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
M = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
paths = sys.argv[3].split(",") if len(sys.argv) > 3 else ["tc", "ffma"]
vo = importlib.import_module("visual-odometry_b200")
synth = importlib.import_module("visual-odometry_b200.synth")
dev = torch.device("cuda:0")
m = synth.nn_map_torch(M, dev)
qn, target = synth.nn_queries_np(Q, M)
q = torch.from_numpy(qn).to(dev)
import threading
class Sampler:
    """SM clock, power and throttle reasons every 20 ms while a filter runs"""
    def __init__(self):
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0)
        self.clk, self.pw, self.reasons = [], [], set()
        self.stop = threading.Event()
    def run(self):
        nv = self.nv
        while not self.stop.is_set():
            self.clk.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.pw.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for name, bit in (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal", 0x20), ("hw_thermal", 0x40)):
                if r & bit:
                    self.reasons.add(name)
            self.stop.wait(0.02)
    def __enter__(self):
        self.t = threading.Thread(target=self.run, daemon=True); self.t.start(); return self
    def __exit__(self, *a):
        self.stop.set(); self.t.join()
    def summary(self):
        return {"sm_mhz_median": float(np.median(self.clk)), "sm_mhz_min": int(min(self.clk)),
                "power_w_median": float(np.median(self.pw)), "power_w_max": float(max(self.pw)),
                "reasons": sorted(self.reasons), "samples": len(self.clk)}
res = {}
for path in paths:
    os.environ["VO_NN_FORCE_PATH"] = path
    nn = vo.NNIndex(0)
    nn.set_stream(torch.cuda.current_stream().cuda_stream)
    nn.set_map_device(m.data_ptr(), M, 11, 1)
    idx = torch.empty(Q, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ts = []
    smp = Sampler()
    smp.__enter__()
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nn.best_match_device(q.data_ptr(), Q, 11, 0.1, idx.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    smp.__exit__()
    got = idx.cpu().numpy()
    cls = np.arange(Q) % 4
    ok = bool(np.array_equal(got[cls < 3], target[cls < 3]) and np.all(got[cls == 3] == -1))
    res[path] = {"ms": ts, "planted_ok": ok, "launches": nn.last_launches(), "rescans": nn.last_rescans(),
                 "queries_per_s": Q / (min(ts[1:]) * 1e-3), "clocks": smp.summary()}
    print(json.dumps({path: res[path]}), flush=True)
    nn.close()

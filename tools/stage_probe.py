"""End-to-end time of vo_triangulate through the host-pointer ABI (pageable numpy buffers in and out)
under different staging settings.   python tools/stage_probe.py            (spawns itself per setting)"""
import importlib, json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    vo = importlib.import_module("visual-odometry_b200")
    synth = importlib.import_module("visual-odometry_b200.synth")
    tv = synth.two_view_problem(int(os.environ.get("N", "9000000")), seed=4, noise=0.2)
    args = (tv["K"], tv["X"], tv["corr"], tv["p1"], tv["p2"])
    for _ in range(2):
        out = vo.triangulate_points(*args)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        out = vo.triangulate_points(*args)
        ts.append(time.perf_counter() - t0)
    n = len(tv["corr"])
    up, down = 8 * n + 8 * len(tv["p1"]) + 8 * len(tv["p2"]), 20 * len(out[0])
    print(json.dumps({"threads": os.environ.get("VO_STAGE_THREADS", "default"), "chunk_mb": os.environ.get("VO_STAGE_CHUNK_MB", "8"),
                      "n_corr": n, "ms_best": round(min(ts) * 1e3, 2), "ms_median": round(sorted(ts)[2] * 1e3, 2),
                      "GBps_total": round((up + down) / min(ts) / 1e9, 1)}))
else:
    for thr, chunk in (("default", "8"), ("16", "8"), ("4", "8"), ("default", "2"), ("default", "4"), ("default", "16"), ("16", "4"), ("12", "16")):
        env = dict(os.environ, VO_STAGE_CHUNK_MB=chunk)
        if thr != "default":
            env["VO_STAGE_THREADS"] = thr
        subprocess.run([sys.executable, __file__, "child"], env=env, check=False)

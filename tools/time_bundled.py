#!/usr/bin/env python
"""Wall clock of host/bin/vo_complete on the bundled sequence under a few environment variants."""
import os, subprocess, sys, tarfile, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(ROOT, "visual-odometry_b200", "host", "bin", "vo_complete")
with tempfile.TemporaryDirectory() as tmp:
    tarfile.open(os.path.join(ROOT, "tests", "golden", "example_data.tar.gz")).extractall(tmp)
    data = [d for d, _, f in os.walk(tmp) if "camera.dat" in f][0]
    for var in ({}, {"VO_PICP_FORCE_STREAM": "1"}, {"CUDA_MODULE_LOADING": "EAGER"}):
        for _ in range(2):
            t0 = time.perf_counter()
            subprocess.run([exe, data], cwd=tmp, env=dict(os.environ, **var), stdout=subprocess.DEVNULL, check=True)
            print(var, f"{time.perf_counter() - t0:.3f} s", flush=True)

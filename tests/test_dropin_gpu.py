"""GPU: the reference's own mains — src/apps/vo_complete.cpp, src/tests/picp_solver_test.cpp,
src/tests/essential_picp_test.cpp, compiled UNCHANGED against the drop-in headers
(visual-odometry_b200/host) and linked to libvo_b200.so by host/build_dropin.sh — run on the GPU
and reproduce what the unmodified CPU reference produces."""
import os
import re
import subprocess
import tarfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(ROOT, "visual-odometry_b200", "host", "bin")


def _need(exe):
    path = os.path.join(BIN, exe)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (host/build_dropin.sh needs the reference checkout)")
    return path


def _matrices(text, label):
    """the 3x3 printed after `label`"""
    tail = text.split(label, 1)[1]
    nums = re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", tail)
    return np.array([float(x) for x in nums[:9]]).reshape(3, 3)


def _loaded_libs(env):
    return env


def test_picp_test_main_recovers_pose(tmp_path):
    exe = _need("picp_test")
    env = dict(os.environ, VO_B200_SEED="5")
    out = subprocess.run([exe], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    R_est, R_gt = _matrices(out.stdout, "R estimated:"), _matrices(out.stdout, "R gt:")
    assert np.allclose(R_est, R_gt, atol=2e-4), out.stdout
    t_est = np.array(re.findall(r"t_est:\s*(\S+)\s+(\S+)\s+(\S+)", out.stdout)[0], dtype=float)
    t_gt = np.array(re.findall(r"t_gt:\s*(\S+)\s+(\S+)\s+(\S+)", out.stdout)[0], dtype=float)
    assert np.allclose(t_est, t_gt, atol=5e-4), out.stdout


def test_whole_test_main(tmp_path):
    exe = _need("whole_test")
    ok = 0
    for seed in ("11", "12", "13"):
        env = dict(os.environ, VO_B200_SEED=seed)
        out = subprocess.run([exe], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        picp = out.stdout.split("PICP RESULTS")[1]
        R_est, R_gt = _matrices(picp, "R estimated:"), _matrices(picp, "R gt:")
        # monocular: rotation is recovered, translation only up to scale (the main prints ratios)
        ok += bool(np.allclose(R_est, R_gt, atol=5e-3))
        assert os.path.exists(os.path.join(tmp_path, "world_triang.txt"))
    assert ok >= 2  # the reference generator sometimes leaves too few common points to converge


def test_vo_complete_on_bundled_data_matches_cpu_reference(tmp_path):
    exe = _need("vo_complete")
    with tarfile.open(os.path.join(HERE, "golden", "example_data.tar.gz")) as tf:
        tf.extractall(tmp_path)
    out = subprocess.run([exe, os.path.join(tmp_path, "data")], cwd=tmp_path, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    ref = np.load(os.path.join(HERE, "golden", "ref_vo_complete_outputs.npz"))
    traj = np.loadtxt(os.path.join(tmp_path, "trajectory_est_complete.txt"))
    gmap = np.loadtxt(os.path.join(tmp_path, "map.txt"))
    assert traj.shape == ref["trajectory_est_complete"].shape == (121, 3)
    # identical data association => identical map size
    assert gmap.shape == ref["map"].shape
    # The pipeline chains 120 relative poses, each from 100 Gauss-Newton rounds on points
    # triangulated with the previous pose: rounding differences grow exponentially along the
    # sequence.  The yardstick is the reference itself: its own sources rebuilt with FMA
    # contraction drift from its SSE2 build by `cpu_dev` (2e-8 at frame 1, 1e-4 at frame 10,
    # 1.6e-2 at frame 60, 0.40 at frame 120).  The GPU run must stay within the same envelope.
    ref_traj = ref["trajectory_est_complete"]
    cpu_dev = np.abs(ref["trajectory_est_complete_fma_build"] - ref_traj).max(1)
    gpu_dev = np.abs(traj - ref_traj).max(1)
    envelope = 10.0 * np.maximum.accumulate(cpu_dev) + 1e-4
    bad = np.nonzero(gpu_dev > envelope)[0]
    assert bad.size == 0, (bad[:5], gpu_dev[bad[:5]], envelope[bad[:5]])
    assert gpu_dev[:10].max() <= 2e-3


@pytest.mark.parametrize("mode", ["pipeline", "classes"])
def test_synthetic_sequence_matches_cpu_reference(tmp_path, mode):
    """Config 5 at test size: the same driver source, built against the drop-in layer (GPU) and
    against the reference's own sources (CPU), on the same synthetic frames.  The GPU build runs
    either through the device-resident frame pipeline (vo_pipe_*) or through the drop-in classes.
    Relative poses must agree per frame to FP32 summation-order noise while the robot drives
    straight (the pipeline is chaotic once it turns: see test_vo_complete_on_bundled_data)."""
    gpu = _need("vo_sequence")
    ref = os.path.join(ROOT, "oracle", "_ref", "bin", "vo_sequence")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/bin/vo_sequence not built")
    import json

    res = {}
    for name, exe in (("ref", ref), ("gpu", gpu)):
        env = dict(os.environ, VO_SEQ_MODE=mode, VO_SEQ_MAPDUMP=str(tmp_path / f"{name}_map.txt"))
        out = subprocess.run([exe, "3000", "24", "1000", "100", str(tmp_path / f"{name}.txt")],
                             capture_output=True, text=True, check=True, env=env).stdout
        res[name] = json.loads(out.strip().splitlines()[-1])
    a, b = np.loadtxt(tmp_path / "ref.txt"), np.loadtxt(tmp_path / "gpu.txt")
    assert a.shape == b.shape == (22, 12)
    assert res["gpu"]["impl"] == ("b200-pipeline" if mode == "pipeline" else "b200")
    # identical data association and map bookkeeping
    for k in ("mean_measurements", "mean_correspondences", "map_points"):
        assert res["gpu"][k] == res["ref"][k], k
    assert np.abs(a - b).max() <= 5e-4, np.abs(a - b).max(1)
    assert res["gpu"]["rot_err_mean_rad"] < 1e-3
    # the maps hold the same landmarks in the same order; low-parallax points are ill-conditioned
    # (their depth amplifies the 1e-5 pose differences), so compare the bulk, not the worst point
    ma, mb = np.loadtxt(tmp_path / "ref_map.txt"), np.loadtxt(tmp_path / "gpu_map.txt")
    assert ma.shape == mb.shape
    err = np.linalg.norm(ma - mb, axis=1) / np.maximum(1.0, np.linalg.norm(ma, axis=1))
    assert np.median(err) <= 5e-4 and np.percentile(err, 90) <= 5e-3, (np.median(err), err.max())

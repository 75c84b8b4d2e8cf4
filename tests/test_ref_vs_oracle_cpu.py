"""CPU: the C restatement (oracle/vo_oracle.c) against the REFERENCE'S OWN code (oracle/_ref,
built from /root/reference + mini_eigen).  This is what pins the oracle's control flow — index
roles, strict '<', compaction order, robust kernel, pose update — to the reference.  (Eigen's
arithmetic ORDER is itself restated in mini_eigen, see DESIGN.md 'Parity status'.)"""
import os
import subprocess

import numpy as np
import pytest

import ref_lib

pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built")
HERE = os.path.dirname(os.path.abspath(__file__))


def test_nn_best_match_bit_exact(oracle, synth):
    m = synth.nn_map_rows_np(0, 3000)
    q, _ = synth.nn_queries_np(500, 3000)
    m[100] = m[7]
    q[3] = m[7]
    for norm in (0.1, 0.9, 50.0):
        ri, rd = ref_lib.nn_best_match(m, q, norm)
        oi, od = oracle.nn_best_match(m, q, norm)
        assert np.array_equal(ri, oi)
        assert np.array_equal(rd[ri >= 0], od[ri >= 0])
        assert np.array_equal(ref_lib.nn_best_match_inplace(m, q, norm), oi)


def test_nn_radius_search_exact(oracle):
    rng = np.random.RandomState(4)
    m = rng.uniform(-0.3, 0.3, (800, 11)).astype(np.float32)
    q = rng.uniform(-0.3, 0.3, (9, 11)).astype(np.float32)
    rc, rl = ref_lib.nn_radius_search(m, q, 0.6, 800)
    oc, ol = oracle.nn_radius_search(m, q, 0.6, 800)
    assert np.array_equal(rc, oc) and np.array_equal(rl, ol)


def test_kdtree_full_agrees_with_brute_force(oracle):
    """bestMatchFull is exact within the radius (eigen_kdtree.h:90-115) except on exact ties,
    where it prefers the right child; bestMatchFast is approximate."""
    b = np.load(os.path.join(HERE, "golden", "bundled_frames.npz"))
    frames = list(b["frames"])
    for a, c in zip(frames[:-1], frames[1:]):
        if c != a + 1:
            continue
        m = np.concatenate([np.zeros((len(b[f"app_{a}"]), 1), np.float32), b[f"app_{a}"]], 1)
        q = np.concatenate([np.zeros((len(b[f"app_{c}"]), 1), np.float32), b[f"app_{c}"]], 1)
        full = ref_lib.kdtree_best_match(m, q, 0.1, leaf=10, full=True)
        oi, _ = oracle.nn_best_match(m, q, 0.1)
        assert np.array_equal(full, oi)
    rng = np.random.RandomState(1)
    m = rng.uniform(-1, 1, (20000, 11)).astype(np.float32)
    q = m[rng.choice(20000, 2000)] + rng.uniform(-0.01, 0.01, (2000, 11)).astype(np.float32)
    full = ref_lib.kdtree_best_match(m, q, 0.1, leaf=10, full=True)
    fast = ref_lib.kdtree_best_match(m, q, 0.1, leaf=10, full=False)
    oi, _ = oracle.nn_best_match(m, q, 0.1)
    assert np.array_equal(full, oi)
    assert np.mean(fast == oi) > 0.5  # approximate: a query near a split plane can miss


def test_project_points_bit_exact(oracle, synth):
    rng = np.random.RandomState(6)
    pts = synth.generate_points3d(rng, 20000)
    T = synth.generate_isometry3f(rng)
    K = synth.default_K()
    for keep in (True, False):
        r, rn = ref_lib.project_points(480, 640, 0, 10, K, T, pts, keep)
        o, on = oracle.project_points(oracle.make_camera(480, 640, 0, 10, K, T), pts, keep)
        assert rn == on and np.array_equal(r, o)


@pytest.mark.parametrize("keep,frac,thr", [(False, 0.0, 10000.0), (False, 0.2, 100.0), (True, 0.2, 100.0)])
def test_picp_rounds(oracle, synth, keep, frac, thr):
    pr = synth.picp_problem(3000, seed=11, outlier_frac=frac, shuffle=True)
    ref = ref_lib.Picp(pr["rows"], pr["cols"], pr["z_near"], pr["z_far"], pr["K"], np.eye(4),
                       pr["world"], pr["image"], thr)
    cam = oracle.make_camera(pr["rows"], pr["cols"], pr["z_near"], pr["z_far"], pr["K"], np.eye(4))
    o = oracle.PicpOracle(cam, pr["world"], pr["image"], thr=thr)
    for r in range(8):
        assert ref.one_round(pr["pairs"], keep) == o.one_round(pr["pairs"], keep) == 1
        s = ref.state()
        assert s["n_in"] == o.st.num_inliers
        scale = np.abs(s["H"]).max()
        # same algorithm, same summation order over correspondences; the 6x6 LDLT inner products
        # may round differently, hence a tolerance instead of equality on the pose
        assert np.max(np.abs(s["H"] - o.H())) <= 2e-6 * scale
        assert abs(s["chi_in"] - o.st.chi_inliers) <= 1e-5 * max(1.0, abs(s["chi_in"]))
        assert abs(s["chi_out"] - o.st.chi_outliers) <= 1e-5 * max(1.0, abs(s["chi_out"]))
        assert np.max(np.abs(s["T"] - o.pose())) <= 2e-5
    if frac == 0.0:
        assert np.allclose(ref.state()["T"], pr["T_gt"], atol=3e-4)


def test_triangulate_bit_exact(oracle, synth):
    tv = synth.two_view_problem(5000, seed=13, noise=0.3)
    app = np.random.RandomState(1).uniform(-1, 1, (len(tv["p2"]), 10)).astype(np.float32)
    corr = tv["corr"][np.random.RandomState(2).permutation(len(tv["corr"]))]
    for a in (None, app):
        rp, rc, ra = ref_lib.triangulate_points(tv["K"], tv["X"], corr, tv["p1"], tv["p2"], a)
        op, oc, oa, _ = oracle.triangulate_points(tv["K"], tv["X"], corr, tv["p1"], tv["p2"], a)
        assert np.array_equal(rc, oc)
        assert np.array_equal(rp, op)
        if a is not None:
            assert np.array_equal(ra, oa)
    X = tv["X"].copy()
    X[:3, 3] *= -1
    rp, rc, _ = ref_lib.triangulate_points(tv["K"], X, corr, tv["p1"], tv["p2"])
    op, oc, _, _ = oracle.triangulate_points(tv["K"], X, corr, tv["p1"], tv["p2"])
    assert len(rp) < len(corr) and np.array_equal(rc, oc) and np.array_equal(rp, op)


def test_reference_vo_complete_runs_on_bundled_data(tmp_path):
    """config 1 sanity anchor: the unmodified reference pipeline on example_data (only where the
    reference checkout exists, i.e. the authoring container)."""
    data = "/root/reference/example_data/data"
    exe = os.path.join(ref_lib.BIN, "vo_complete")
    if not (os.path.isdir(data) and os.path.exists(exe)):
        pytest.skip("reference data / binary not present")
    subprocess.run([exe, data], cwd=tmp_path, check=True, stdout=subprocess.DEVNULL, timeout=120)
    out = subprocess.run([os.path.join(ref_lib.BIN, "evaluation"), data], cwd=tmp_path, check=True,
                         capture_output=True, text=True, timeout=120).stdout
    ratio = float([l for l in out.splitlines() if "ratio" in l][0].split(":")[1])
    assert abs(ratio - 0.47337) < 0.02  # README.md:77
    traj = np.loadtxt(os.path.join(tmp_path, "trajectory_est_complete.txt"))
    assert traj.shape == (121, 3)

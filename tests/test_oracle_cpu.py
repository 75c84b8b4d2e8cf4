"""CPU tests: the oracle against known answers (no GPU).  The reference ships no golden vectors
(SURVEY.md §4); what pins the oracle here is (i) the bundled dataset's landmark ids, (ii)
closed-form answers (planted matches, exact ties, LDLT vs numpy, noise-free geometry with a
known ground truth), (iii) float64 restatements."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def bundled():
    return np.load(os.path.join(HERE, "golden", "bundled_frames.npz"))


def _vec11(app):
    return np.concatenate([np.arange(len(app), dtype=np.float32)[:, None], app], axis=1)


def id_truth(ids_map, ids_q):
    """compute_corr.cpp:68-82: a query's true match is the map row with the same landmark id."""
    pos = {int(v): i for i, v in enumerate(ids_map)}
    return np.array([pos.get(int(v), -1) for v in ids_q], dtype=np.int32)


def test_sqdist_order_matches_eigen_sse2_redux(oracle):
    rng = np.random.RandomState(0)
    for _ in range(200):
        p = rng.uniform(-1, 1, 10).astype(np.float32)
        q = rng.uniform(-1, 1, 10).astype(np.float32)
        s = ((p - q) * (p - q)).astype(np.float32)
        a = s[0:4] + s[4:8]                      # two packets added lane-wise
        r = np.float32(np.float32(a[0] + a[2]) + np.float32(a[1] + a[3]))  # movehl + add_ss
        r = np.float32(np.float32(r + s[8]) + s[9])                         # scalar tail
        assert oracle.sqdist(p, q) == r


def test_sqdist_small_dims(oracle):
    rng = np.random.RandomState(1)
    for dim in (1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 17, 31):
        p = rng.uniform(-1, 1, dim).astype(np.float32)
        q = rng.uniform(-1, 1, dim).astype(np.float32)
        exact = float(np.sum((p.astype(np.float64) - q.astype(np.float64)) ** 2))
        assert abs(oracle.sqdist(p, q) - exact) <= 1e-6 * max(1.0, exact)


def test_nn_bundled_matches_id_truth(oracle, bundled):
    frames = list(bundled["frames"])
    for a, b in zip(frames[:-1], frames[1:]):
        if b != a + 1:
            continue
        m, q = _vec11(bundled[f"app_{a}"]), _vec11(bundled[f"app_{b}"])
        idx, d2 = oracle.nn_best_match(m, q, 0.1)
        truth = id_truth(bundled[f"ids_{a}"], bundled[f"ids_{b}"])
        assert np.array_equal(idx, truth)
        assert np.all(d2[truth >= 0] == 0.0)  # appearances are noise-free


def test_nn_bundled_vs_world(oracle, bundled):
    w = _vec11(bundled["world_app"])
    for k in bundled["frames"]:
        idx, _ = oracle.nn_best_match(w, _vec11(bundled[f"app_{k}"]), 0.1)
        assert np.array_equal(bundled["world_ids"][idx], bundled[f"ids_{k}"])


def test_nn_ties_none_and_radius_edge(oracle):
    rng = np.random.RandomState(3)
    m = rng.uniform(-1, 1, (64, 11)).astype(np.float32)
    m[40] = m[7]       # duplicate rows: lowest index wins (strict '<')
    m[41] = m[7]
    q = m[[7, 40, 41]].copy()
    idx, _ = oracle.nn_best_match(m, q, 0.1)
    assert idx.tolist() == [7, 7, 7]
    far = np.full((1, 11), 5.0, np.float32)
    assert oracle.nn_best_match(m, far, 0.1)[0].tolist() == [-1]
    # exactly on the radius: d2 == norm*norm must NOT match (strict '<')
    one = np.zeros((1, 11), np.float32)
    qq = np.zeros((1, 11), np.float32)
    qq[0, 1] = 0.5
    assert oracle.nn_best_match(one, qq, 0.5)[0].tolist() == [-1]
    assert oracle.nn_best_match(one, qq, np.nextafter(np.float32(0.5), np.float32(1)))[0].tolist() == [0]
    # empty map
    assert oracle.nn_best_match(np.zeros((0, 11), np.float32), qq, 0.5)[0].tolist() == [-1]


def test_nn_radius_search(oracle):
    rng = np.random.RandomState(4)
    m = rng.uniform(-0.3, 0.3, (500, 11)).astype(np.float32)
    q = rng.uniform(-0.3, 0.3, (5, 11)).astype(np.float32)
    counts, lst = oracle.nn_radius_search(m, q, 0.6, 500)
    d = ((m[None, :, 1:].astype(np.float64) - q[:, None, 1:]) ** 2).sum(-1)
    for i in range(5):
        ref = np.nonzero(d[i] < 0.36 - 1e-6)[0]
        got = lst[i, :counts[i]]
        assert set(ref).issubset(set(got.tolist()))
        assert np.all(np.diff(got) > 0)


def test_ldlt_matches_numpy(oracle):
    rng = np.random.RandomState(5)
    for n in (2, 6):
        for _ in range(50):
            A = rng.normal(size=(n, n))
            A = (A @ A.T + np.eye(n) * 0.1).astype(np.float32)
            b = rng.normal(size=n).astype(np.float32)
            x = oracle.ldlt_solve(A, b)
            ref = np.linalg.solve(A.astype(np.float64), b.astype(np.float64))
            assert np.allclose(x, ref, rtol=2e-3, atol=2e-4)
    # singular 2x2 (parallel rays): Eigen's LDLT zeroes the dead pivot's component
    A = np.array([[1, 1], [1, 1]], np.float32)
    x = oracle.ldlt_solve(A, np.array([1, 1], np.float32))
    assert np.all(np.isfinite(x))


def test_v2t_euler(oracle):
    v = np.array([0.1, -0.2, 0.3, 0.05, -0.07, 0.11], np.float32)
    T = oracle.v2t_euler(v)
    cx, sx, cy, sy, cz, sz = np.cos(v[3]), np.sin(v[3]), np.cos(v[4]), np.sin(v[4]), np.cos(v[5]), np.sin(v[5])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    assert np.allclose(T[:3, :3], Rx @ Ry @ Rz, atol=1e-6)
    assert np.allclose(T[:3, 3], v[:3])


def test_project_points(oracle, synth):
    rng = np.random.RandomState(6)
    pts = synth.generate_points3d(rng, 5000)
    T = synth.generate_isometry3f(rng)
    K = synth.default_K()
    cam = oracle.make_camera(480, 640, 0, 10, K, T)
    uv, n_in = oracle.project_points(cam, pts, True)
    ref, ok = synth.project_np(K, T, pts)
    assert uv.shape == (5000, 2)
    # borderline points may flip between float32 and float64; the bulk must agree
    ok32 = uv[:, 0] != -1
    assert np.sum(ok32 != ok) <= 2
    both = ok32 & ok
    assert np.allclose(uv[both], ref[both], atol=2e-3)
    uvc, n_in2 = oracle.project_points(cam, pts, False)
    assert n_in2 == n_in == ok32.sum() == len(uvc)
    assert np.array_equal(uvc, uv[ok32])


def test_picp_recovers_pose_and_matches_f64(oracle, synth):
    """picp_solver_test.cpp:45-78 shape: start at identity, converge to the hidden pose."""
    pr = synth.picp_problem(2000, seed=11)
    cam = oracle.make_camera(pr["rows"], pr["cols"], pr["z_near"], pr["z_far"], pr["K"], np.eye(4))
    o = oracle.PicpOracle(cam, pr["world"], pr["image"], thr=10000.0)
    for r in range(30):
        assert o.one_round(pr["pairs"]) == 1
        o.one_round_f64(pr["pairs"])
        if r == 0:
            H32, H64 = o.H(), o.H64m()
            assert np.allclose(H32, H64, rtol=1e-4, atol=1e-6 * np.abs(H64).max())
            assert np.allclose(H32, H32.T)
    assert np.allclose(o.pose(), pr["T_gt"], atol=2e-4)
    assert np.allclose(o.pose64(), pr["T_gt"].astype(np.float64), atol=2e-5)
    assert o.st.num_inliers == len(pr["pairs"])


def test_picp_outliers_and_keep(oracle, synth):
    pr = synth.picp_problem(1500, seed=12, outlier_frac=0.2)
    cam = oracle.make_camera(pr["rows"], pr["cols"], pr["z_near"], pr["z_far"], pr["K"], np.eye(4))
    for keep in (False, True):
        o = oracle.PicpOracle(cam, pr["world"], pr["image"], thr=100.0)
        o.one_round(pr["pairs"], keep)
        o.one_round_f64(pr["pairs"], keep)
        assert o.st.chi_outliers > 0 and 0 < o.st.num_inliers < len(pr["pairs"])
        assert abs(o.st.chi_inliers - o.stats64[0]) <= 1e-4 * o.stats64[0]
        assert o.st.num_inliers == int(o.stats64[2])
        assert np.allclose(o.H(), o.H64m(), rtol=1e-3, atol=1e-5 * np.abs(o.H64m()).max())


def test_triangulation_noise_free(oracle, synth):
    """essential_picp_test.cpp:77-82: noise-free two-view triangulation returns the GT points."""
    tv = synth.two_view_problem(3000, seed=13)
    pts, cn, _, src = oracle.triangulate_points(tv["K"], tv["X"], tv["corr"], tv["p1"], tv["p2"])
    assert len(pts) >= 0.95 * len(tv["corr"])
    gt = tv["points"][tv["corr"][src, 0]]
    assert np.allclose(pts, gt, atol=5e-3)
    assert np.array_equal(cn[:, 1], np.arange(len(pts)))
    assert np.array_equal(cn[:, 0], tv["corr"][src, 1])


def test_triangulation_rejects_behind(oracle, synth):
    tv = synth.two_view_problem(500, seed=14)
    # swap the two pixels of some correspondences: rays diverge -> negative ray parameter
    p1, p2 = tv["p1"].copy(), tv["p2"].copy()
    X = tv["X"].copy()
    X[:3, 3] *= -1.0  # wrong baseline sign: most points end up behind a camera
    pts, _, _, src = oracle.triangulate_points(tv["K"], X, tv["corr"], p1, p2)
    assert len(pts) < len(tv["corr"])
    assert np.all(np.diff(src) > 0)  # order-preserving compaction


def test_synth_hash_numpy_torch_identical(synth):
    import torch

    a = synth.nn_map_rows_np(1000, 1300)
    b = synth.nn_map_rows_torch(1000, 1300, "cpu").numpy()
    assert np.array_equal(a, b)
    assert a[:, 1:].min() >= -1 and a[:, 1:].max() < 1
    q, target = synth.nn_queries_np(64, 5000)
    m = synth.nn_map_rows_np(0, 5000)
    exact = (np.arange(64) % 4) < 2
    assert np.array_equal(q[exact, 1:], m[target[exact], 1:])


def test_device_ldlt_source_is_bit_identical_to_oracle_on_host(tmp_path):
    """linalg.cuh's register-resident LDL^T (what the PICP and triangulation kernels run) is
    plain C++: compile it for the host without FMA contraction and compare it bit for bit with
    oracle_ldlt_solve on SPD, indefinite, singular and badly scaled systems."""
    import subprocess

    root = os.path.dirname(HERE)
    exe = str(tmp_path / "ldlt_host_check")
    subprocess.check_call(["make", "-C", os.path.join(root, "oracle")], stdout=subprocess.DEVNULL)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17",
                           "-I", os.path.join(root, "visual-odometry_b200", "csrc"),
                           os.path.join(HERE, "ldlt_host_check.cpp"),
                           os.path.join(root, "oracle", "libvo_oracle.so"),
                           "-Wl,-rpath," + os.path.join(root, "oracle"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "bad6=0 bad2=0" in out.stdout

"""GPU parity: vo_triangulate / vo_project_points vs the oracle (1e-5 relative, flags equal)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("n", [1, 31, 1024, 1025, 10000, 300000])
def test_triangulate_vs_oracle(vo, oracle, synth, n):
    tv = synth.two_view_problem(n, seed=31, noise=0.3)
    app = np.random.RandomState(1).uniform(-1, 1, (len(tv["p2"]), 10)).astype(np.float32)
    pts, cn, oa, src = vo.triangulate_points(tv["K"], tv["X"], tv["corr"], tv["p1"], tv["p2"],
                                             appearances2=app, want_src=True)
    opts, ocn, ooa, osrc = oracle.triangulate_points(tv["K"], tv["X"], tv["corr"], tv["p1"],
                                                     tv["p2"], app)
    # compare per ORIGINAL correspondence (a borderline sign flip would shift every later index)
    common, ia, ib = np.intersect1d(src, osrc, return_indices=True)
    flips = (len(src) - len(common)) + (len(osrc) - len(common))
    assert flips <= max(1, n // 100000), f"{flips} accept/reject mismatches"
    scale = max(np.abs(opts).max(), 1.0) if len(opts) else 1.0
    assert np.max(np.abs(pts[ia] - opts[ib]), initial=0.0) <= 20 * TOL * scale
    if flips == 0:
        assert np.array_equal(src, osrc)
        assert np.array_equal(cn, ocn)
        assert np.array_equal(oa, ooa)
    assert np.all(np.diff(src) > 0)
    assert np.array_equal(cn[:, 1], np.arange(len(cn)))
    assert np.array_equal(cn[:, 0], tv["corr"][src, 1])
    assert np.array_equal(oa, app[cn[:, 0]])


def test_triangulate_noise_free_recovers_gt(vo, synth):
    tv = synth.two_view_problem(50000, seed=32)
    pts, cn, src = vo.triangulate_points(tv["K"], tv["X"], tv["corr"], tv["p1"], tv["p2"],
                                         want_src=True)
    assert len(pts) >= 0.95 * len(tv["corr"])
    assert np.allclose(pts, tv["points"][tv["corr"][src, 0]], atol=5e-3)


def test_triangulate_mostly_rejected_and_shuffled(vo, oracle, synth):
    tv = synth.two_view_problem(20000, seed=33)
    X = tv["X"].copy()
    X[:3, 3] *= -1.0
    corr = tv["corr"][np.random.RandomState(2).permutation(len(tv["corr"]))]
    pts, cn, src = vo.triangulate_points(tv["K"], X, corr, tv["p1"], tv["p2"], want_src=True)
    opts, ocn, _, osrc = oracle.triangulate_points(tv["K"], X, corr, tv["p1"], tv["p2"])
    assert len(osrc) < len(corr)
    assert abs(len(src) - len(osrc)) <= 1
    if len(src) == len(osrc):
        assert np.array_equal(src, osrc) and np.array_equal(cn, ocn)


def test_triangulate_empty_and_bad_index(vo):
    K, X = np.eye(3, dtype=np.float32), np.eye(4, dtype=np.float32)
    pts, cn = vo.triangulate_points(K, X, np.zeros((0, 2), np.int32), np.zeros((4, 2), np.float32),
                                    np.zeros((4, 2), np.float32))
    assert pts.shape == (0, 3) and cn.shape == (0, 2)
    with pytest.raises(vo.VoError):
        vo.triangulate_points(K, X, np.array([[0, 9]], np.int32), np.zeros((4, 2), np.float32),
                              np.zeros((4, 2), np.float32))


@pytest.mark.parametrize("n", [1, 1000, 123457])
@pytest.mark.parametrize("keep", [True, False])
def test_project_points_vs_oracle(vo, oracle, synth, n, keep):
    rng = np.random.RandomState(n)
    pts = synth.generate_points3d(rng, n)
    T = synth.generate_isometry3f(rng, 0.3)
    K = synth.default_K()
    cam = vo.Camera(480, 640, 0, 10, K, T)
    uv, n_in = cam.projectPoints(pts, keep)
    ouv, on_in = oracle.project_points(oracle.make_camera(480, 640, 0, 10, K, T), pts, keep)
    assert abs(n_in - on_in) <= 1
    if n_in == on_in:
        assert uv.shape == ouv.shape
        assert np.allclose(uv, ouv, rtol=TOL, atol=640 * TOL)

"""GPU: the device-resident frame pipeline (include/vo_b200.h §5, csrc/pipeline.cu).
 * vo_pipe_merge_cloud == PointCloudVector::update (reference include/PointCloud.h:52-66): first
   equal appearance is overwritten, else append in order; appended points take part in later
   matches; float equality (-0.0 == +0.0, NaN equals nothing).
 * the whole loop against the CPU reference is in tests/test_dropin_gpu.py."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_update(map_pts, map_app, pts, app):
    """the reference's sequential semantics, restated"""
    for p, a in zip(pts, app):
        hit = None
        for j, b in enumerate(map_app):
            if np.all(b == a):  # float ==: NaN never equal, -0.0 == 0.0
                hit = j
                break
        if hit is None:
            map_pts.append(p.copy())
            map_app.append(a.copy())
        else:
            map_pts[hit] = p.copy()


def _pipe(vo, max_pts=4096, max_map=5000):
    abi = vo._abi if hasattr(vo, "_abi") else __import__("importlib").import_module("visual-odometry_b200._abi")
    lib = vo.lib()
    cam = abi.vo_camera()
    cam.rows, cam.cols, cam.z_near, cam.z_far = 480, 640, 0, 5
    K = np.array([[180, 0, 320], [0, 180, 240], [0, 0, 1]], np.float32)
    for j in range(3):
        for i in range(3):
            cam.K[j * 3 + i] = K[i, j]
    for i in range(16):
        cam.T[i] = 1.0 if i % 5 == 0 else 0.0
    h = C.c_void_p()
    assert lib.vo_pipe_create(C.byref(h), 0, C.byref(cam), max_pts, max_map) == 0, lib.vo_last_error()
    return lib, h


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_map_merge_matches_sequential_semantics(vo):
    lib, h = _pipe(vo)
    rng = np.random.RandomState(3)
    pool = rng.uniform(-1, 1, (300, 10)).astype(np.float32)
    pool[10] = pool[3]                      # exact duplicate appearance
    pool[20, 4], pool[21] = 0.0, pool[20]   # +0 / -0 twins
    pool[21, 4] = -0.0
    pool[30, 7] = np.nan                    # equals nothing, not even itself
    ref_pts, ref_app = [], []
    for call in range(25):
        n = int(rng.randint(1, 400))
        ids = rng.randint(0, 300, n)
        if call % 5 == 0:
            ids[: n // 2] = ids[n // 2: n // 2 + n // 2][: n // 2]  # many repeats inside one call
        app = pool[ids].copy()
        pts = rng.uniform(-5, 5, (n, 3)).astype(np.float32)
        th = 0.1 * call
        X = np.eye(4, dtype=np.float32)
        X[:3, :3] = [[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]]
        X[:3, 3] = [0.3 * call, -0.1, 0.05]
        Xc = np.ascontiguousarray(X.T).reshape(-1)  # column-major
        assert lib.vo_pipe_merge_cloud(h, _p(pts), _p(app), n, Xc.ctypes.data_as(C.POINTER(C.c_float))) == 0
        moved = (pts.astype(np.float64) @ X[:3, :3].T.astype(np.float64) + X[:3, 3]).astype(np.float32)
        reference_update(ref_pts, ref_app, moved, app)
    cap = 6000
    got_pts = np.zeros((cap, 3), np.float32)
    got_app = np.zeros((cap, 10), np.float32)
    n_map = C.c_int64(0)
    assert lib.vo_pipe_get_map(h, _p(got_pts), _p(got_app), cap, C.byref(n_map)) == 0
    lib.vo_pipe_destroy(h)
    assert n_map.value == len(ref_app)
    ra, rp = np.array(ref_app), np.array(ref_pts)
    ga, gp = got_app[: n_map.value], got_pts[: n_map.value]
    assert np.array_equal(np.isnan(ga), np.isnan(ra))
    assert np.array_equal(np.nan_to_num(ga), np.nan_to_num(ra))  # same appearances in the same order
    assert np.abs(gp - rp).max() <= 1e-5 * max(1.0, np.abs(rp).max())


def test_map_overflow_is_reported_not_fatal(vo):
    lib, h = _pipe(vo, max_pts=1024, max_map=100)
    rng = np.random.RandomState(4)
    app = rng.uniform(-1, 1, (300, 10)).astype(np.float32)
    pts = rng.uniform(-1, 1, (300, 3)).astype(np.float32)
    X = np.eye(4, dtype=np.float32).reshape(-1)
    assert lib.vo_pipe_merge_cloud(h, _p(pts), _p(app), 300, X.ctypes.data_as(C.POINTER(C.c_float))) == 0
    n_map = C.c_int64(0)
    assert lib.vo_pipe_get_map(h, None, None, 0, C.byref(n_map)) == 0
    assert n_map.value == 100
    # a dropped claim must not cut later keys off their probe sequence.  The 100 stored
    # points are still found after the overflow: merging them again (moved) changes no count and
    # overwrites their positions in place.
    Xf = C.POINTER(C.c_float)
    moved = pts[:100] + np.float32(7.0)
    assert lib.vo_pipe_merge_cloud(h, _p(moved), _p(app[:100].copy()), 100, X.ctypes.data_as(Xf)) == 0
    got_pts, got_app = np.zeros((100, 3), np.float32), np.zeros((100, 10), np.float32)
    assert lib.vo_pipe_get_map(h, _p(got_pts), _p(got_app), 100, C.byref(n_map)) == 0
    assert n_map.value == 100
    assert np.array_equal(got_app, app[:100]) and np.array_equal(got_pts, moved)
    # ... and a new sequence on the same handle starts from an empty map
    f = np.zeros((4, 2), np.float32)
    a = rng.uniform(-1, 1, (4, 10)).astype(np.float32)
    assert lib.vo_pipe_first_frame(h, _p(f), _p(a), 4) == 0
    assert lib.vo_pipe_get_map(h, None, None, 0, C.byref(n_map)) == 0
    assert n_map.value == 0
    assert lib.vo_pipe_merge_cloud(h, _p(pts), _p(app), 50, X.ctypes.data_as(Xf)) == 0
    assert lib.vo_pipe_get_map(h, None, None, 0, C.byref(n_map)) == 0
    assert n_map.value == 50
    lib.vo_pipe_destroy(h)


def test_call_order_is_enforced(vo):
    lib, h = _pipe(vo)
    res = (C.c_byte * 256)()
    pts = np.zeros((4, 2), np.float32)
    app = np.zeros((4, 10), np.float32)
    assert lib.vo_pipe_step(h, _p(pts), _p(app), 4, 10, C.c_float(1e4), C.byref(res)) == -4  # VO_ERR_STATE
    X = np.eye(4, dtype=np.float32).reshape(-1)
    assert lib.vo_pipe_bootstrap(h, X.ctypes.data_as(C.POINTER(C.c_float))) == -4
    lib.vo_pipe_destroy(h)


def test_degenerate_frames_do_not_break_the_loop(vo):
    """empty and unmatched frames: no matches -> no correspondences -> the pose stays the identity
    and nothing is triangulated; the loop keeps running"""
    lib, h = _pipe(vo, max_pts=2048, max_map=4000)
    abi = __import__("importlib").import_module("visual-odometry_b200._abi")
    rng = np.random.RandomState(9)

    def frame(n):
        return (rng.uniform(0, 600, (n, 2)).astype(np.float32), rng.uniform(-1, 1, (n, 10)).astype(np.float32))

    class Result(C.Structure):
        _fields_ = [("T", C.c_float * 16), ("n_measurements", C.c_int64), ("n_matches", C.c_int64),
                    ("n_correspondences", C.c_int64), ("map_points", C.c_int64),
                    ("chi_inliers", C.c_float), ("n_inliers", C.c_int32), ("map_overflow", C.c_int32)]

    p0, a0 = frame(50)
    assert lib.vo_pipe_first_frame(h, _p(p0), _p(a0), 50) == 0
    p1, a1 = frame(40)  # unrelated appearances: no match
    corr = np.zeros((64, 2), np.int32)
    n = C.c_int64(-1)
    assert lib.vo_pipe_second_frame(h, _p(p1), _p(a1), 40, _p(corr), 64, C.byref(n)) == 0
    assert n.value == 0
    X = np.eye(4, dtype=np.float32).reshape(-1)
    assert lib.vo_pipe_bootstrap(h, X.ctypes.data_as(C.POINTER(C.c_float))) == 0
    res = Result()
    for size in (30, 0, 0, 25, 1):
        pts, app = frame(max(size, 1))
        assert lib.vo_pipe_step(h, _p(pts), _p(app), size, 5, C.c_float(1e4), C.byref(res)) == 0, lib.vo_last_error()
        assert res.n_measurements == size and res.n_matches == 0 and res.n_correspondences == 0
        assert np.allclose(np.array(res.T[:]).reshape(4, 4), np.eye(4), atol=1e-6)
    # a frame that repeats the previous one exactly: every point matches itself
    pts, app = frame(64)
    assert lib.vo_pipe_step(h, _p(pts), _p(app), 64, 5, C.c_float(1e4), C.byref(res)) == 0
    assert lib.vo_pipe_step(h, _p(pts), _p(app), 64, 5, C.c_float(1e4), C.byref(res)) == 0
    assert res.n_matches == 64
    n_map = C.c_int64(-1)
    assert lib.vo_pipe_get_map(h, None, None, 0, C.byref(n_map)) == 0
    assert n_map.value >= 0
    lib.vo_pipe_destroy(h)


def test_python_frame_pipeline_on_a_synthetic_pair_sequence(vo, synth):
    """vo.FramePipeline (the ctypes mirror of vo_pipe_*): a static scene seen from a camera that
    moves along its optical axis; matches are the identity, the pose is recovered up to the
    monocular scale, and the map holds every landmark once."""
    rng = np.random.RandomState(21)
    K = np.array([[180, 0, 320], [0, 180, 240], [0, 0, 1]], np.float64)
    n = 600
    world = np.stack([rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(2.0, 4.5, n)], 1)
    app = rng.uniform(-1, 1, (n, 10)).astype(np.float32)

    def view(tz):  # camera at (0,0,tz): world-in-camera translation (0,0,-tz)
        pc = world - np.array([0, 0, tz])
        uv = (pc @ K.T)[:, :2] / pc[:, 2:3]
        return uv.astype(np.float32)

    cam = vo.Camera(480, 640, 0, 50, K, np.eye(4))
    pipe = vo.FramePipeline(cam, max_points_per_frame=1024, max_map_points=2000)
    pipe.first_frame(view(0.0), app)
    matches = pipe.second_frame(view(0.1), app)
    assert np.array_equal(matches, np.stack([np.arange(n), np.arange(n)], 1))
    X = np.eye(4)
    X[2, 3] = -0.1  # ground truth relative pose of the first pair
    pipe.bootstrap(X)
    for k in range(2, 6):
        pose, info = pipe.step(view(0.1 * k), app, rounds=30)
        assert info["n_matches"] == n and info["n_correspondences"] == n
        assert np.allclose(pose[:3, :3], np.eye(3), atol=2e-4)
        assert np.allclose(pose[:3, 3], [0, 0, -0.1], atol=2e-3)
    pts, apps = pipe.map()
    assert len(pts) == n and np.array_equal(apps, app)
    pipe.close()


def test_cluster_join_equals_single_cta_join(vo):
    """frames of >= 2048 measurements take assoc_join_cluster_kernel (8 CTAs, counts exchanged through
    distributed shared memory): same matches, same correspondences, same poses, bit for bit, as the
    one-CTA kernel (VO_PIPE_JOIN_SINGLE=1, read once per process: run in a child).  The scene has
    unmatched measurements on both sides and landmarks that appear and disappear, so hits and joins
    are proper subsets in every slice."""
    import json, subprocess, sys, textwrap
    code = textwrap.dedent("""
        import importlib, json, sys
        import numpy as np
        sys.path.insert(0, %r)
        vo = importlib.import_module("visual-odometry_b200")
        rng = np.random.RandomState(5)
        K = np.array([[180, 0, 320], [0, 180, 240], [0, 0, 1]], np.float64)
        n = 5000
        world = np.stack([rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(2.0, 4.5, n)], 1)
        app = rng.uniform(-1, 1, (n, 10)).astype(np.float32)
        def view(k):
            keep = rng.uniform(size=n) > 0.2                      # a fifth of the landmarks missing per frame
            pc = world[keep] - np.array([0, 0, 0.1 * k])
            uv = ((pc @ K.T)[:, :2] / pc[:, 2:3]).astype(np.float32)
            a = app[keep].copy()
            extra = 300                                           # clutter that matches nothing
            uv = np.concatenate([uv, rng.uniform(0, 400, (extra, 2)).astype(np.float32)])
            a = np.concatenate([a, rng.uniform(-1, 1, (extra, 10)).astype(np.float32)])
            return uv, a
        cam = vo.Camera(480, 640, 0, 50, K, np.eye(4))
        pipe = vo.FramePipeline(cam, max_points_per_frame=8192, max_map_points=20000)
        pipe.first_frame(*view(0))
        m = pipe.second_frame(*view(1))
        X = np.eye(4); X[2, 3] = -0.1
        pipe.bootstrap(X)
        out = {"matches": np.asarray(m).tolist(), "poses": [], "info": []}
        for k in range(2, 7):
            pose, info = pipe.step(*view(k), rounds=20)
            out["poses"].append(np.asarray(pose, np.float32).view(np.uint32).tolist())
            out["info"].append([int(info["n_matches"]), int(info["n_correspondences"])])
        pts, apps = pipe.map()
        out["map"] = [len(pts), int(np.asarray(pts, np.float32).view(np.uint32).sum() %% (1 << 31))]
        pipe.close()
        print(json.dumps(out))
    """ % ROOT)
    res = {}
    for mode in ("cluster", "single"):
        env = dict(os.environ)
        env.pop("VO_PIPE_JOIN_SINGLE", None)
        if mode == "single":
            env["VO_PIPE_JOIN_SINGLE"] = "1"
        o = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout
        res[mode] = json.loads(o.strip().splitlines()[-1])
    assert res["cluster"]["info"][0][0] > 3000 and res["cluster"]["info"][0][1] > 2000
    assert res["cluster"] == res["single"]

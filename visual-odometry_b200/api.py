"""Host-side mirror of the reference call surface, over the C ABI (numpy in, numpy out).

Names and argument meaning follow the reference (file:line relative to its checkout):
  Camera                  include/camera.h:16-62, src/camera.cpp
  PICPSolver              include/picp_solver.h:18-79, src/picp_solver.cpp
  bruteForceBestMatch     include/brute_force_search.h:22-41
  bruteForceSearch        include/brute_force_search.h:3-20
  triangulate_points      src/utils.cpp:51-134
  FramePipeline           the loop body of src/apps/vo_complete.cpp:150-178 (device-resident)
Everything computes on the GPU through libvo_b200.so; nothing here does arithmetic on the host.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import check, lib, vo_camera, vo_picp_state


def _f32(a, shape_last=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape_last is not None and (a.ndim != 2 or a.shape[1] != shape_last):
        raise ValueError(f"expected an (N,{shape_last}) float32 array, got {a.shape}")
    return a


def _pairs(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    if a.size == 0:
        return a.reshape(0, 2)
    if a.ndim != 2 or a.shape[1] != 2:
        raise ValueError(f"expected an (N,2) int32 array of index pairs, got {a.shape}")
    return a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class Camera:
    """Pinhole camera (camera.h:16-62).  K is 3x3, world_in_camera_pose is 4x4 (row/col as in
    numpy; converted to Eigen's column-major at the ABI)."""

    def __init__(self, rows=100, cols=100, z_near=0, z_far=10, camera_matrix=None,
                 world_in_camera_pose=None, device=0):
        self._rows, self._cols = int(rows), int(cols)
        self._z_near, self._z_far = int(z_near), int(z_far)
        self._K = np.eye(3, dtype=np.float32) if camera_matrix is None else \
            np.array(camera_matrix, dtype=np.float32).reshape(3, 3)
        self._T = np.eye(4, dtype=np.float32) if world_in_camera_pose is None else \
            np.array(world_in_camera_pose, dtype=np.float32).reshape(4, 4)
        self.device = device

    def rows(self):
        return self._rows

    def cols(self):
        return self._cols

    def cameraMatrix(self):
        return self._K

    def worldInCameraPose(self):
        return self._T

    def setWorldInCameraPose(self, pose):
        self._T = np.array(pose, dtype=np.float32).reshape(4, 4)

    def to_struct(self):
        cam = vo_camera()
        cam.rows, cam.cols, cam.z_near, cam.z_far = self._rows, self._cols, self._z_near, self._z_far
        cam.K[:] = self._K.T.reshape(-1).tolist()  # column-major
        cam.T[:] = self._T.T.reshape(-1).tolist()
        return cam

    def projectPoints(self, world_points, keep_indices=False):
        """camera.cpp:16-37 -> (image_points, num_points_inside)."""
        return project_points(self, world_points, keep_indices)


def project_points(camera, world_points, keep_indices=False):
    w = _f32(world_points, 3)
    out = np.empty((w.shape[0], 2), dtype=np.float32)
    n_out, n_in = C.c_int64(0), C.c_int64(0)
    cam = camera.to_struct()
    check(lib().vo_project_points(camera.device, C.byref(cam), _ptr(w), w.shape[0],
                                  1 if keep_indices else 0, _ptr(out), C.byref(n_out),
                                  C.byref(n_in)), "vo_project_points")
    return out[: n_out.value], int(n_in.value)


class NNIndex:
    """A resident appearance map answering bruteForceBestMatch / bruteForceSearch queries.

    Rows are `row_stride` floats with the first `skip_cols` ignored (Vector11f = [id | 10-D
    appearance]: row_stride 11, skip_cols 1 — defs.h:7, vo_complete.cpp:22)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().vo_nn_create(C.byref(self._h), device), "vo_nn_create")
        self.device = device
        self.row_stride = 0
        self.skip_cols = 0

    def close(self):
        if self._h:
            lib().vo_nn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        check(lib().vo_nn_set_stream(self._h, C.c_void_p(cuda_stream)), "vo_nn_set_stream")

    def synchronize(self):
        check(lib().vo_nn_synchronize(self._h), "vo_nn_synchronize")

    def set_map(self, rows, skip_cols=1):
        rows = _f32(rows)
        if rows.ndim != 2:
            raise ValueError("map must be (M, row_stride)")
        self.row_stride, self.skip_cols = rows.shape[1], skip_cols
        check(lib().vo_nn_set_map(self._h, _ptr(rows), rows.shape[0], rows.shape[1], skip_cols),
              "vo_nn_set_map")

    def set_map_device(self, dev_ptr, n_rows, row_stride, skip_cols=1):
        self.row_stride, self.skip_cols = row_stride, skip_cols
        check(lib().vo_nn_set_map_device(self._h, C.c_void_p(dev_ptr), n_rows, row_stride,
                                         skip_cols), "vo_nn_set_map_device")

    def best_match(self, queries, norm, want_d2=False):
        q = _f32(queries)
        idx = np.empty(q.shape[0], dtype=np.int32)
        d2 = np.empty(q.shape[0], dtype=np.float32) if want_d2 else None
        check(lib().vo_nn_best_match(self._h, _ptr(q), q.shape[0], q.shape[1], float(norm),
                                     _ptr(idx), _ptr(d2) if want_d2 else None), "vo_nn_best_match")
        return (idx, d2) if want_d2 else idx

    def best_match_device(self, q_ptr, n_queries, query_stride, norm, idx_ptr, d2_ptr=0):
        check(lib().vo_nn_best_match_device(self._h, C.c_void_p(q_ptr), n_queries, query_stride,
                                            float(norm), C.c_void_p(idx_ptr),
                                            C.c_void_p(d2_ptr) if d2_ptr else None),
              "vo_nn_best_match_device")

    def last_launches(self):
        """[(queries per thread, threads, query tiles, map splits), ...] of the last best_match."""
        n = C.c_int(0)
        buf = np.zeros((16, 4), dtype=np.int32)
        check(lib().vo_nn_last_launches(self._h, _ptr(buf), 16, C.byref(n)), "vo_nn_last_launches")
        return [tuple(int(x) for x in row) for row in buf[: min(n.value, 16)]]

    def last_rescans(self):
        """blocks the tensor-core filter handed to the exact re-rank in the last best_match (-1: the
        FP32 filter ran)"""
        n = C.c_int64(0)
        check(lib().vo_nn_last_rescans(self._h, C.byref(n)), "vo_nn_last_rescans")
        return int(n.value)

    def radius_search(self, queries, norm, max_per_query=0):
        q = _f32(queries)
        counts = np.empty(q.shape[0], dtype=np.int32)
        lst = np.full((q.shape[0], max_per_query), -1, dtype=np.int32) if max_per_query else None
        check(lib().vo_nn_radius_search(self._h, _ptr(q), q.shape[0], q.shape[1], float(norm),
                                        _ptr(counts), _ptr(lst) if max_per_query else None,
                                        max_per_query), "vo_nn_radius_search")
        return counts, lst


class ShardedNN:
    """bruteForceBestMatch on several GPUs of one node from ONE process (vo_comm_*): map replicated,
    queries sharded in contiguous blocks, indices gathered with one ncclAllGather."""

    def __init__(self, n_gpus):
        self._c = C.c_void_p()
        check(lib().vo_comm_init_all(C.byref(self._c), int(n_gpus)), "vo_comm_init_all")

    def close(self):
        if self._c:
            lib().vo_comm_destroy(self._c)
            self._c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        return int(lib().vo_comm_size(self._c))

    def set_map(self, rows, skip_cols=1):
        rows = _f32(rows)
        check(lib().vo_nn_set_map_replicated(self._c, _ptr(rows), rows.shape[0], rows.shape[1], skip_cols),
              "vo_nn_set_map_replicated")

    def best_match(self, queries, norm):
        q = _f32(queries)
        idx = np.empty(q.shape[0], dtype=np.int32)
        check(lib().vo_nn_best_match_sharded(self._c, _ptr(q), q.shape[0], q.shape[1], float(norm), _ptr(idx)),
              "vo_nn_best_match_sharded")
        return idx


def bruteForceBestMatch(points, query, norm, device=0):
    """brute_force_search.h:22-41 for one query (or a batch): index of the best row or -1."""
    points = _f32(points)
    q = _f32(query)
    single = q.ndim == 1
    q = q.reshape(1, -1) if single else q
    nn = NNIndex(device)
    try:
        nn.set_map(points, skip_cols=1)
        idx = nn.best_match(q, norm)
    finally:
        nn.close()
    return int(idx[0]) if single else idx


def bruteForceSearch(points, query, norm, device=0):
    """brute_force_search.h:3-20: indices (ascending) of all rows within `norm` of the query."""
    points = _f32(points)
    q = _f32(query).reshape(1, -1)
    nn = NNIndex(device)
    try:
        nn.set_map(points, skip_cols=1)
        counts, _ = nn.radius_search(q, norm, 0)
        m = int(counts[0])
        if m == 0:
            return np.empty(0, dtype=np.int32)
        _, lst = nn.radius_search(q, norm, m)
    finally:
        nn.close()
    return lst[0]


class PICPSolver:
    """picp_solver.h:18-79.  Usage mirrors the reference: init(camera, world, image), then
    oneRound(correspondences, keep_outliers) repeatedly; compute() runs many rounds without
    leaving the device."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().vo_picp_create(C.byref(self._h), device), "vo_picp_create")
        self.device = device
        self._kernel_threshold = 1000.0  # picp_solver.cpp:13
        self._damping = 1.0              # :10
        self._min_num_inliers = 0        # :11
        self._camera = None
        self._corr_key = None

    def close(self):
        if self._h:
            lib().vo_picp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _push_params(self):
        check(lib().vo_picp_set_params(self._h, self._kernel_threshold, self._damping,
                                       self._min_num_inliers), "vo_picp_set_params")

    def kernelThreshold(self):
        return self._kernel_threshold

    def setKernelThreshold(self, thr):
        self._kernel_threshold = float(thr)
        self._push_params()

    def set_stream(self, cuda_stream):
        check(lib().vo_picp_set_stream(self._h, C.c_void_p(cuda_stream)), "vo_picp_set_stream")

    def synchronize(self):
        check(lib().vo_picp_synchronize(self._h), "vo_picp_synchronize")

    def init(self, camera, world_points, image_points):
        w, im = _f32(world_points, 3), _f32(image_points, 2)
        self._camera = camera
        cam = camera.to_struct()
        self._push_params()
        check(lib().vo_picp_init(self._h, C.byref(cam), _ptr(w), w.shape[0], _ptr(im), im.shape[0]),
              "vo_picp_init")
        self._corr_key = None

    def init_device(self, camera, world_ptr, n_world, image_ptr, n_image):
        self._camera = camera
        cam = camera.to_struct()
        self._push_params()
        check(lib().vo_picp_init_device(self._h, C.byref(cam), C.c_void_p(world_ptr), n_world,
                                        C.c_void_p(image_ptr), n_image), "vo_picp_init_device")
        self._corr_key = None

    def set_correspondences(self, correspondences):
        p = _pairs(correspondences)
        check(lib().vo_picp_set_correspondences(self._h, _ptr(p), p.shape[0]),
              "vo_picp_set_correspondences")

    def set_correspondences_device(self, pairs_ptr, n_pairs):
        check(lib().vo_picp_set_correspondences_device(self._h, C.c_void_p(pairs_ptr), n_pairs),
              "vo_picp_set_correspondences_device")

    def compute(self, keep_outliers=False, rounds=1):
        check(lib().vo_picp_compute(self._h, 1 if keep_outliers else 0, int(rounds)),
              "vo_picp_compute")

    def oneRound(self, correspondences, keep_outliers):
        """picp_solver.cpp:98-112.  Returns the reference's bool."""
        p = _pairs(correspondences)
        check(lib().vo_picp_one_round(self._h, _ptr(p), p.shape[0], 1 if keep_outliers else 0),
              "vo_picp_one_round")
        return bool(self.state().last_ok)

    def state(self):
        st = vo_picp_state()
        check(lib().vo_picp_get_state(self._h, C.byref(st)), "vo_picp_get_state")
        return st

    # accessors (picp_solver.h:41-50); each one synchronises
    def pose(self):
        return np.array(self.state().T[:], dtype=np.float32).reshape(4, 4).T.copy()

    def camera(self):
        cam = Camera(self._camera.rows(), self._camera.cols(), self._camera._z_near,
                     self._camera._z_far, self._camera.cameraMatrix(), self.pose(), self.device)
        return cam

    def H(self):
        return np.array(self.state().H[:], dtype=np.float32).reshape(6, 6).T.copy()

    def b(self):
        return np.array(self.state().b[:], dtype=np.float32)

    def chiInliers(self):
        return float(self.state().chi_inliers)

    def chiOutliers(self):
        return float(self.state().chi_outliers)

    def numInliers(self):
        return int(self.state().num_inliers)


def triangulate_points(k, X, correspondences, p1_img, p2_img, appearances2=None, device=0,
                       want_src=False):
    """utils.cpp:51-134.  Returns (triangulated (n,3), correspondences_new (n,2)
    [, appearances (n,10)] [, src (n,)])."""
    K = np.array(k, dtype=np.float32).reshape(3, 3).T.copy().reshape(-1)
    Xc = np.array(X, dtype=np.float32).reshape(4, 4).T.copy().reshape(-1)
    corr = _pairs(correspondences)
    p1, p2 = _f32(p1_img, 2), _f32(p2_img, 2)
    n = corr.shape[0]
    app = _f32(appearances2, 10) if appearances2 is not None else None
    pts = np.empty((n, 3), dtype=np.float32)
    cn = np.empty((n, 2), dtype=np.int32)
    oa = np.empty((n, 10), dtype=np.float32) if app is not None else None
    src = np.empty(n, dtype=np.int32) if want_src else None
    ns = C.c_int64(0)
    check(lib().vo_triangulate(device, K.ctypes.data_as(_abi.c_f32p), Xc.ctypes.data_as(_abi.c_f32p),
                               _ptr(corr), n, _ptr(p1), p1.shape[0], _ptr(p2), p2.shape[0],
                               _ptr(app) if app is not None else None, _ptr(pts), _ptr(cn),
                               _ptr(oa) if oa is not None else None,
                               _ptr(src) if want_src else None, C.byref(ns)), "vo_triangulate")
    m = ns.value
    out = [pts[:m], cn[:m]]
    if app is not None:
        out.append(oa[:m])
    if want_src:
        out.append(src[:m])
    return tuple(out)


class vo_pipe_result(C.Structure):
    """include/vo_b200.h: vo_pipe_result"""
    _fields_ = [("T", C.c_float * 16), ("n_measurements", C.c_int64), ("n_matches", C.c_int64),
                ("n_correspondences", C.c_int64), ("map_points", C.c_int64),
                ("chi_inliers", C.c_float), ("n_inliers", C.c_int32), ("map_overflow", C.c_int32)]


class FramePipeline:
    """The loop body of the reference's main (src/apps/vo_complete.cpp:150-178) with the frames, the
    match lists, the triangulated cloud and the map resident on the device (vo_pipe_*).

        pipe = FramePipeline(camera)
        pipe.first_frame(points0, appearances0)
        matches = pipe.second_frame(points1, appearances1)   # (n,2) int32: for the epipolar init
        pipe.bootstrap(X)                                    # X from estimate_transform (host)
        pose, info = pipe.step(points, appearances)          # every further frame
        map_points, map_appearances = pipe.map()
    """

    def __init__(self, camera, device=0, max_points_per_frame=32768, max_map_points=1 << 20):
        self._h = C.c_void_p()
        cam = camera.to_struct()
        check(lib().vo_pipe_create(C.byref(self._h), device, C.byref(cam), max_points_per_frame,
                                   max_map_points), "vo_pipe_create")
        self._max_map = max_map_points

    def close(self):
        if self._h:
            lib().vo_pipe_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _frame(points, appearances):
        p, a = _f32(points, 2), _f32(appearances, 10)
        if p.shape[0] != a.shape[0]:
            raise ValueError("points and appearances must have the same length")
        return p, a

    def first_frame(self, points, appearances):
        p, a = self._frame(points, appearances)
        check(lib().vo_pipe_first_frame(self._h, _ptr(p), _ptr(a), p.shape[0]), "vo_pipe_first_frame")
        self._n0 = p.shape[0]

    def second_frame(self, points, appearances):
        p, a = self._frame(points, appearances)
        cap = max(1, min(self._n0, p.shape[0]))
        corr = np.empty((cap, 2), dtype=np.int32)
        n = C.c_int64(0)
        check(lib().vo_pipe_second_frame(self._h, _ptr(p), _ptr(a), p.shape[0], _ptr(corr), cap,
                                         C.byref(n)), "vo_pipe_second_frame")
        return corr[: n.value].copy()

    def bootstrap(self, X):
        Xc = np.array(X, dtype=np.float32).reshape(4, 4).T.copy().reshape(-1)
        check(lib().vo_pipe_bootstrap(self._h, Xc.ctypes.data_as(_abi.c_f32p)), "vo_pipe_bootstrap")

    def step(self, points, appearances, rounds=100, kernel_threshold=10000.0):
        p, a = self._frame(points, appearances)
        res = vo_pipe_result()
        check(lib().vo_pipe_step(self._h, _ptr(p), _ptr(a), p.shape[0], rounds,
                                 C.c_float(kernel_threshold), C.byref(res)), "vo_pipe_step")
        pose = np.array(res.T[:], dtype=np.float32).reshape(4, 4).T.copy()
        info = {k: getattr(res, k) for k in ("n_measurements", "n_matches", "n_correspondences",
                                             "map_points", "chi_inliers", "n_inliers", "map_overflow")}
        return pose, info

    def merge_cloud(self, points, appearances, X=np.eye(4)):
        """PointCloudVector::update (PointCloud.h:52-66) of a host cloud moved by X"""
        p, a = _f32(points, 3), _f32(appearances, 10)
        Xc = np.array(X, dtype=np.float32).reshape(4, 4).T.copy().reshape(-1)
        check(lib().vo_pipe_merge_cloud(self._h, _ptr(p), _ptr(a), p.shape[0],
                                        Xc.ctypes.data_as(_abi.c_f32p)), "vo_pipe_merge_cloud")

    def map(self):
        n = C.c_int64(0)
        check(lib().vo_pipe_get_map(self._h, None, None, 0, C.byref(n)), "vo_pipe_get_map")
        pts = np.empty((max(n.value, 1), 3), dtype=np.float32)
        app = np.empty((max(n.value, 1), 10), dtype=np.float32)
        check(lib().vo_pipe_get_map(self._h, _ptr(pts), _ptr(app), pts.shape[0], C.byref(n)),
              "vo_pipe_get_map")
        return pts[: n.value], app[: n.value]

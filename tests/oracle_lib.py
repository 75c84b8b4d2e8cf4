"""ctypes binding of oracle/libvo_oracle.so — the CPU restatement used as the parity checker.
TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "libvo_oracle.so")


class oracle_camera(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("z_near", C.c_int32),
                ("z_far", C.c_int32), ("K", C.c_float * 9), ("T", C.c_float * 16)]


class oracle_picp_state(C.Structure):
    _fields_ = [("T", C.c_float * 16), ("H", C.c_float * 36), ("b", C.c_float * 6),
                ("chi_inliers", C.c_float), ("chi_outliers", C.c_float),
                ("num_inliers", C.c_int32), ("rounds_done", C.c_int32), ("last_ok", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        L = C.CDLL(_SO)
        L.oracle_sqdist.restype = C.c_float
        L.oracle_sqdist.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_nn_best_match.restype = None
        L.oracle_nn_best_match.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                           C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        L.oracle_nn_radius_search.restype = None
        L.oracle_nn_radius_search.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                              C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                              C.c_int32]
        L.oracle_project_points.restype = None
        L.oracle_project_points.argtypes = [C.POINTER(oracle_camera), C.c_void_p, C.c_int64, C.c_int,
                                            C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.oracle_picp_one_round.restype = C.c_int
        L.oracle_picp_one_round.argtypes = [C.POINTER(oracle_picp_state), C.POINTER(oracle_camera),
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                            C.c_float, C.c_float, C.c_int32]
        L.oracle_picp_one_round_f64.restype = C.c_int
        L.oracle_picp_one_round_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.POINTER(oracle_camera), C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_int64, C.c_int, C.c_double,
                                                C.c_double]
        L.oracle_triangulate_points.restype = C.c_int64
        L.oracle_triangulate_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_ldlt_solve.restype = None
        L.oracle_ldlt_solve.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_v2t_euler.restype = None
        L.oracle_v2t_euler.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def make_camera(rows, cols, z_near, z_far, K, T):
    cam = oracle_camera()
    cam.rows, cam.cols, cam.z_near, cam.z_far = int(rows), int(cols), int(z_near), int(z_far)
    cam.K[:] = np.asarray(K, dtype=np.float32).reshape(3, 3).T.reshape(-1).tolist()
    cam.T[:] = np.asarray(T, dtype=np.float32).reshape(4, 4).T.reshape(-1).tolist()
    return cam


def sqdist(p, q):
    p, q = _f32(p), _f32(q)
    return float(lib().oracle_sqdist(_p(p), _p(q), p.size))


def nn_best_match(rows, queries, norm, skip_cols=1):
    rows, queries = _f32(rows), _f32(queries)
    idx = np.empty(queries.shape[0], dtype=np.int32)
    d2 = np.empty(queries.shape[0], dtype=np.float32)
    lib().oracle_nn_best_match(_p(rows), rows.shape[0], rows.shape[1], skip_cols, _p(queries),
                               queries.shape[0], queries.shape[1], float(norm), _p(idx), _p(d2))
    return idx, d2


def nn_radius_search(rows, queries, norm, max_per_query, skip_cols=1):
    rows, queries = _f32(rows), _f32(queries)
    counts = np.empty(queries.shape[0], dtype=np.int32)
    lst = np.full((queries.shape[0], max(max_per_query, 1)), -1, dtype=np.int32)
    lib().oracle_nn_radius_search(_p(rows), rows.shape[0], rows.shape[1], skip_cols, _p(queries),
                                  queries.shape[0], queries.shape[1], float(norm), _p(counts),
                                  _p(lst) if max_per_query else None, max_per_query)
    return counts, lst[:, :max_per_query]


def project_points(cam, world, keep_indices):
    world = _f32(world)
    out = np.empty((world.shape[0], 2), dtype=np.float32)
    n_out, n_in = C.c_int64(0), C.c_int64(0)
    lib().oracle_project_points(C.byref(cam), _p(world), world.shape[0], 1 if keep_indices else 0,
                                _p(out), C.byref(n_out), C.byref(n_in))
    return out[: n_out.value], int(n_in.value)


class PicpOracle:
    """Sequential-FP32 PICPSolver (the reference's arithmetic) + an FP64 twin."""

    def __init__(self, cam, world, image, thr=1000.0, damping=1.0, min_inliers=0):
        self.cam = cam
        self.world, self.image = _f32(world), _f32(image)
        self.thr, self.damping, self.min_inliers = thr, damping, min_inliers
        self.st = oracle_picp_state()
        self.st.T[:] = list(cam.T)
        self.T64 = np.array(list(cam.T), dtype=np.float64)
        self.H64 = np.zeros(36)
        self.b64 = np.zeros(6)
        self.stats64 = np.zeros(3)

    def one_round(self, pairs, keep_outliers=False):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32)
        return lib().oracle_picp_one_round(C.byref(self.st), C.byref(self.cam), _p(self.world),
                                           _p(self.image), _p(pairs), pairs.shape[0],
                                           1 if keep_outliers else 0, self.thr, self.damping,
                                           self.min_inliers)

    def one_round_f64(self, pairs, keep_outliers=False):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32)
        return lib().oracle_picp_one_round_f64(_p(self.T64), _p(self.H64), _p(self.b64),
                                               _p(self.stats64), C.byref(self.cam), _p(self.world),
                                               _p(self.image), _p(pairs), pairs.shape[0],
                                               1 if keep_outliers else 0, float(self.thr),
                                               float(self.damping))

    def pose(self):
        return np.array(self.st.T[:], dtype=np.float32).reshape(4, 4).T.copy()

    def H(self):
        return np.array(self.st.H[:], dtype=np.float32).reshape(6, 6).T.copy()

    def b(self):
        return np.array(self.st.b[:], dtype=np.float32)

    def pose64(self):
        return self.T64.reshape(4, 4).T.copy()

    def H64m(self):
        return self.H64.reshape(6, 6).T.copy()


def triangulate_points(K, X, corr, p1, p2, app2=None):
    Kc = np.asarray(K, dtype=np.float32).reshape(3, 3).T.copy().reshape(-1)
    Xc = np.asarray(X, dtype=np.float32).reshape(4, 4).T.copy().reshape(-1)
    corr = np.ascontiguousarray(corr, dtype=np.int32)
    p1, p2 = _f32(p1), _f32(p2)
    n = corr.shape[0]
    pts = np.empty((n, 3), dtype=np.float32)
    cn = np.empty((n, 2), dtype=np.int32)
    src = np.empty(n, dtype=np.int32)
    app = _f32(app2) if app2 is not None else None
    oa = np.empty((n, 10), dtype=np.float32) if app is not None else None
    ns = lib().oracle_triangulate_points(_p(Kc), _p(Xc), _p(corr), n, _p(p1), _p(p2), _p(app),
                                         _p(pts), _p(cn), _p(oa), _p(src))
    return pts[:ns], cn[:ns], (oa[:ns] if oa is not None else None), src[:ns]


def ldlt_solve(A, rhs):
    A = np.asarray(A, dtype=np.float32)
    n = A.shape[0]
    Ac = np.ascontiguousarray(A.T).reshape(-1)  # column-major
    rhs = _f32(rhs)
    x = np.empty(n, dtype=np.float32)
    lib().oracle_ldlt_solve(n, _p(Ac), _p(rhs), _p(x))
    return x


def v2t_euler(v):
    v = _f32(v)
    T = np.empty(16, dtype=np.float32)
    lib().oracle_v2t_euler(_p(v), _p(T))
    return T.reshape(4, 4).T.copy()

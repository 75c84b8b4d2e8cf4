"""ctypes binding of oracle/_ref/libvo_ref.so — the reference's OWN sources (compiled from
/root/reference against third_party/mini_eigen by oracle/build_ref.sh).  TEST INFRASTRUCTURE.
Absent => available() is False and the tests that need it skip."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_ref", "libvo_ref.so")
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
_lib = None


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_SO)
        vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float
        L.ref_nn_best_match.argtypes = [vp, i64, vp, i64, f32, vp, vp]
        L.ref_nn_best_match_inplace.argtypes = [vp, i64, vp, i64, f32, vp]
        L.ref_nn_radius_search.argtypes = [vp, i64, vp, i64, f32, vp, vp, i32]
        L.ref_kdtree_best_match.argtypes = [vp, i64, vp, i64, f32, C.c_int, C.c_int, vp]
        L.ref_project_points.argtypes = [C.c_int] * 4 + [vp, vp, vp, i64, C.c_int, vp,
                                                         C.POINTER(i64), C.POINTER(i64)]
        L.ref_picp_create.restype = vp
        L.ref_picp_create.argtypes = [C.c_int] * 4 + [vp, vp, vp, i64, vp, i64, f32]
        L.ref_picp_destroy.argtypes = [vp]
        L.ref_picp_one_round.restype = C.c_int
        L.ref_picp_one_round.argtypes = [vp, vp, i64, C.c_int]
        L.ref_picp_get_state.argtypes = [vp, vp, vp, vp, C.POINTER(f32), C.POINTER(f32),
                                         C.POINTER(i32)]
        L.ref_triangulate_points.restype = i64
        L.ref_triangulate_points.argtypes = [vp, vp, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _cm(M, n):
    return np.ascontiguousarray(np.asarray(M, dtype=np.float32).reshape(n, n).T).reshape(-1)


def nn_best_match(rows, queries, norm):
    rows, queries = _f32(rows), _f32(queries)
    idx = np.empty(len(queries), np.int32)
    d2 = np.empty(len(queries), np.float32)
    lib().ref_nn_best_match(_p(rows), len(rows), _p(queries), len(queries), norm, _p(idx), _p(d2))
    return idx, d2


def nn_best_match_inplace(rows, queries, norm):
    """bruteForceBestMatch over the caller's (M,11) float32 buffer, no copy; thread-safe."""
    assert rows.dtype == np.float32 and rows.flags.c_contiguous and rows.shape[1] == 11
    queries = _f32(queries)
    idx = np.empty(len(queries), np.int32)
    lib().ref_nn_best_match_inplace(_p(rows), len(rows), _p(queries), len(queries), norm, _p(idx))
    return idx


def nn_radius_search(rows, queries, norm, max_per_query):
    rows, queries = _f32(rows), _f32(queries)
    counts = np.empty(len(queries), np.int32)
    lst = np.full((len(queries), max_per_query), -1, np.int32)
    lib().ref_nn_radius_search(_p(rows), len(rows), _p(queries), len(queries), norm, _p(counts),
                               _p(lst), max_per_query)
    return counts, lst


def kdtree_best_match(rows, queries, norm, leaf=10, full=True):
    rows, queries = _f32(rows), _f32(queries)
    idx = np.empty(len(queries), np.int32)
    lib().ref_kdtree_best_match(_p(rows), len(rows), _p(queries), len(queries), norm, leaf,
                                1 if full else 0, _p(idx))
    return idx


def project_points(rows, cols, z_near, z_far, K, T, world, keep_indices):
    world = _f32(world)
    out = np.empty((len(world), 2), np.float32)
    n_out, n_in = C.c_int64(0), C.c_int64(0)
    Kc, Tc = _cm(K, 3), _cm(T, 4)
    lib().ref_project_points(rows, cols, z_near, z_far, _p(Kc), _p(Tc), _p(world), len(world),
                             1 if keep_indices else 0, _p(out), C.byref(n_out), C.byref(n_in))
    return out[: n_out.value], int(n_in.value)


class Picp:
    def __init__(self, rows, cols, z_near, z_far, K, T, world, image, thr):
        self.world, self.image = _f32(world), _f32(image)
        Kc, Tc = _cm(K, 3), _cm(T, 4)
        self.h = lib().ref_picp_create(rows, cols, z_near, z_far, _p(Kc), _p(Tc), _p(self.world),
                                       len(self.world), _p(self.image), len(self.image), thr)

    def one_round(self, pairs, keep=False):
        pairs = np.ascontiguousarray(pairs, np.int32)
        return lib().ref_picp_one_round(self.h, _p(pairs), len(pairs), 1 if keep else 0)

    def state(self):
        T, H, b = np.empty(16, np.float32), np.empty(36, np.float32), np.empty(6, np.float32)
        ci, co, n = C.c_float(0), C.c_float(0), C.c_int32(0)
        lib().ref_picp_get_state(self.h, _p(T), _p(H), _p(b), C.byref(ci), C.byref(co), C.byref(n))
        return dict(T=T.reshape(4, 4).T.copy(), H=H.reshape(6, 6).T.copy(), b=b, chi_in=ci.value,
                    chi_out=co.value, n_in=n.value)

    def __del__(self):
        try:
            lib().ref_picp_destroy(self.h)
        except Exception:
            pass


def triangulate_points(K, X, corr, p1, p2, app2=None):
    corr = np.ascontiguousarray(corr, np.int32)
    p1, p2 = _f32(p1), _f32(p2)
    n = len(corr)
    pts = np.empty((n, 3), np.float32)
    cn = np.empty((n, 2), np.int32)
    app = _f32(app2) if app2 is not None else None
    oa = np.empty((n, 10), np.float32) if app is not None else None
    Kc, Xc = _cm(K, 3), _cm(X, 4)
    ns = lib().ref_triangulate_points(_p(Kc), _p(Xc), _p(corr), n, _p(p1), len(p1), _p(p2), len(p2),
                                      _p(app), _p(pts), _p(cn), _p(oa))
    return pts[:ns], cn[:ns], (oa[:ns] if oa is not None else None)

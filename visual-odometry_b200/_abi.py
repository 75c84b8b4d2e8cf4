"""ctypes binding of include/vo_b200.h (one prototype per exported symbol)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VO_B200_LIB: another build of the same library (kernel experiments, tools/build_variants.sh)
_LIB_PATH = os.environ.get("VO_B200_LIB") or os.path.join(_HERE, "lib", "libvo_b200.so")

c_f32p = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)


class VoError(RuntimeError):
    pass


class vo_camera(C.Structure):
    _fields_ = [
        ("rows", C.c_int32),
        ("cols", C.c_int32),
        ("z_near", C.c_int32),
        ("z_far", C.c_int32),
        ("K", C.c_float * 9),
        ("T", C.c_float * 16),
    ]


class vo_picp_state(C.Structure):
    _fields_ = [
        ("T", C.c_float * 16),
        ("H", C.c_float * 36),
        ("b", C.c_float * 6),
        ("chi_inliers", C.c_float),
        ("chi_outliers", C.c_float),
        ("num_inliers", C.c_int32),
        ("rounds_done", C.c_int32),
        ("last_ok", C.c_int32),
    ]


# name -> (restype, argtypes); must list every symbol declared in include/vo_b200.h
PROTOTYPES = {
    "vo_abi_version": (C.c_int, []),
    "vo_last_error": (C.c_char_p, []),
    "vo_device_count": (C.c_int, []),
    "vo_launch_count": (C.c_int64, []),
    "vo_measure_ffma_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "vo_measure_ffma2_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "vo_nn_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "vo_nn_destroy": (C.c_int, [C.c_void_p]),
    "vo_nn_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vo_nn_synchronize": (C.c_int, [C.c_void_p]),
    "vo_nn_set_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int]),
    "vo_nn_set_map_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int]),
    "vo_nn_best_match": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p],
    ),
    "vo_nn_best_match_device": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p],
    ),
    "vo_nn_last_rescans": (C.c_int, [C.c_void_p, c_i64p]),
    "vo_nn_last_launches": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "vo_nn_radius_search": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int32],
    ),
    "vo_comm_init_all": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "vo_comm_destroy": (C.c_int, [C.c_void_p]),
    "vo_comm_size": (C.c_int, [C.c_void_p]),
    "vo_nn_set_map_replicated": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int]),
    "vo_nn_best_match_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p]),
    "vo_picp_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "vo_picp_destroy": (C.c_int, [C.c_void_p]),
    "vo_picp_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vo_picp_synchronize": (C.c_int, [C.c_void_p]),
    "vo_picp_set_params": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_int32]),
    "vo_picp_init": (
        C.c_int,
        [C.c_void_p, C.POINTER(vo_camera), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64],
    ),
    "vo_picp_init_device": (
        C.c_int,
        [C.c_void_p, C.POINTER(vo_camera), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64],
    ),
    "vo_picp_set_correspondences": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "vo_picp_set_correspondences_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "vo_picp_compute": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "vo_picp_compute_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, c_f32p]),
    "vo_picp_one_round": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "vo_picp_get_state": (C.c_int, [C.c_void_p, C.POINTER(vo_picp_state)]),
    "vo_triangulate": (
        C.c_int,
        [C.c_int, c_f32p, c_f32p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
         C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_i64p],
    ),
    "vo_triangulate_workspace_bytes": (C.c_int64, [C.c_int64]),
    "vo_triangulate_device": (
        C.c_int,
        [C.c_void_p, c_f32p, c_f32p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "vo_triangulate_device_ex": (
        C.c_int,
        [C.c_void_p, c_f32p, c_f32p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "vo_pipe_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(vo_camera), C.c_int64,
                                 C.c_int64]),
    "vo_pipe_destroy": (C.c_int, [C.c_void_p]),
    "vo_pipe_first_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "vo_pipe_second_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                       C.c_int64, c_i64p]),
    "vo_pipe_bootstrap": (C.c_int, [C.c_void_p, c_f32p]),
    "vo_pipe_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float,
                               C.c_void_p]),
    "vo_pipe_merge_cloud": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, c_f32p]),
    "vo_pipe_get_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, c_i64p]),
    "vo_project_points": (
        C.c_int,
        [C.c_int, C.POINTER(vo_camera), C.c_void_p, C.c_int64, C.c_int, C.c_void_p, c_i64p, c_i64p],
    ),
}

_lib = None


def lib_path():
    return _LIB_PATH


def lib():
    """Load libvo_b200.so (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise VoError(
                f"{_LIB_PATH} is missing: run `make` (or __graft_entry__.build()); "
                "there is no CPU fallback"
            )
        handle = C.CDLL(_LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().vo_last_error().decode("utf-8", "replace")
        raise VoError(f"{what} failed with status {rc}: {msg}")


def launch_count():
    return int(lib().vo_launch_count())


def device_count():
    n = lib().vo_device_count()
    if n < 0:
        check(n, "vo_device_count")
    return n


def measure_ffma_peak(device=0, packed=False):
    """FP32 FMA throughput in TFLOP/s: scalar FFMA, or packed FFMA2 (fma.rn.f32x2)."""
    out = C.c_double(0.0)
    fn = lib().vo_measure_ffma2_peak if packed else lib().vo_measure_ffma_peak
    check(fn(device, C.byref(out)), "vo_measure_ffma_peak")
    return out.value

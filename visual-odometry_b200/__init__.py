"""visual-odometry_b200 — B200 (sm_100a) hot path of lucanunz/Visual-odometry.

The product is ``lib/libvo_b200.so`` (hand-written CUDA behind the C ABI in ``include/vo_b200.h``)
plus the C++ drop-in headers under ``host/``.  This Python package is the ctypes binding the
tests and ``bench.py`` use; it mirrors the reference's call surface (``PICPSolver``, ``Camera``,
``bruteForceBestMatch``, ``triangulate_points``) so parity tests read like the reference's mains.

The directory name contains a hyphen, so import it with
``importlib.import_module("visual-odometry_b200")`` (see ``__graft_entry__.py``).

There is NO CPU fallback: every compute entry point raises ``VoError`` when the CUDA library is
missing or no device is visible.
"""
from ._abi import VoError, lib, lib_path, launch_count, device_count, measure_ffma_peak  # noqa: F401
from .api import (  # noqa: F401
    Camera,
    FramePipeline,
    NNIndex,
    PICPSolver,
    ShardedNN,
    bruteForceBestMatch,
    bruteForceSearch,
    project_points,
    triangulate_points,
)

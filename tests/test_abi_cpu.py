"""CPU: the C-ABI library loads and exports every symbol include/vo_b200.h declares; compute
entry points fail loudly (no CPU fallback) when no device is visible."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "vo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vo_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(vo):
    lib = vo.lib()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in vo_b200.h but not exported"
    from importlib import import_module

    abi = import_module("visual-odometry_b200._abi")
    assert sorted(abi.PROTOTYPES) == declared
    assert lib.vo_abi_version() == 1


def test_struct_layouts_match_oracle(vo, oracle):
    from importlib import import_module

    abi = import_module("visual-odometry_b200._abi")
    assert C.sizeof(abi.vo_camera) == C.sizeof(oracle.oracle_camera) == 4 * 4 + 9 * 4 + 16 * 4
    assert C.sizeof(abi.vo_picp_state) == C.sizeof(oracle.oracle_picp_state)


def test_no_cpu_fallback(vo):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; this test checks the no-device behaviour")
    with pytest.raises(vo.VoError):
        vo.NNIndex(0)
    with pytest.raises(vo.VoError):
        vo.PICPSolver(0)
    with pytest.raises(vo.VoError):
        vo.triangulate_points(np.eye(3), np.eye(4), np.zeros((1, 2), np.int32),
                              np.zeros((1, 2), np.float32), np.zeros((1, 2), np.float32))


def test_argument_errors(vo):
    lib = vo.lib()
    assert lib.vo_nn_create(None, 0) == -1  # VO_ERR_ARG
    assert b"null" in lib.vo_last_error()
    assert lib.vo_triangulate_workspace_bytes(-1) < 0
    assert lib.vo_triangulate_workspace_bytes(5000) >= 5 * 8

"""The C-ABI library loads and exports every symbol include/lqgk.h declares (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import pytest

from lqg_b200 import abi
from tests import helpers as H

HEADER = os.path.join(H.ROOT, "include", "lqgk.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lqgk_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ["lqgk_lqr_backward_f32", "lqgk_lqr_backward_f64", "lqgk_kf_forward_f32", "lqgk_kf_forward_f64",
              "lqgk_loglik_fwd_f32", "lqgk_loglik_fwd_f64", "lqgk_loglik_vjp_f32", "lqgk_loglik_vjp_f64",
              "lqgk_pack_obs_f32", "lqgk_pack_obs_f64", "lqgk_workspace_bytes", "lqgk_dims_supported", "lqgk_strerror",
              "lqgk_version"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    if not os.path.exists(abi.LIB_PATH):
        pytest.fail(f"{abi.LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(abi.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_host_side_queries_without_gpu():
    lib = abi.load_library()
    assert "sm_100a" in lib.version()
    d = abi.LqgkDims(64, 20, 100, 2, 2, 1, 2, 2)
    assert lib.lib.lqgk_dims_supported(ctypes.byref(d)) == 1
    bad = abi.LqgkDims(64, 20, 100, 3, 7, 1, 2, 2)
    assert lib.lib.lqgk_dims_supported(ctypes.byref(bad)) == 0
    small, big = lib.workspace_bytes(d, abi.MODE_VJP, 32), lib.workspace_bytes(d, abi.MODE_VJP, 0)
    assert 0 < small < big and lib.workspace_bytes(d, abi.MODE_FWD, 0) < big
    assert "unsupported" in lib.strerror(-2)


def test_struct_layouts_match_header():
    assert ctypes.sizeof(abi.LqgkDims) == 40
    assert ctypes.sizeof(abi.LqgkMat) == 24 and ctypes.sizeof(abi.LqgkSpec) == 12 * 24
    assert ctypes.sizeof(abi.LqgkMatGrad) == 16 and ctypes.sizeof(abi.LqgkSpecGrad) == 8 * 16

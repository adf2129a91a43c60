"""CPU check of the kernel mathematics: the CUDA kernels' own step functions (lqg_b200/csrc/lqgk_core.h,
lqgk_stages.h, lqgk_pack.h) compiled for the host (tests/emul) against the float64 oracle, through the same
C ABI structs the GPU library uses.  The GPU run of the same checks is tests/test_gpu_parity.py."""
import os
import subprocess

import pytest
import torch

from lqg_b200 import abi
from tests import helpers as H


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(H.ROOT, "tests", "emul", "lqgk_emul.cpp")
    if not os.path.exists(H.EMUL_PATH) or os.path.getmtime(H.EMUL_PATH) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ftemplate-depth=2000", "-shared", "-fPIC", "-o", H.EMUL_PATH, src])
    return abi.Library(H.EMUL_PATH)


CPU = torch.device("cpu")


@pytest.mark.parametrize("name,d", [("bounded", 2), ("subjective", 2), ("relobs", 2), ("bounded2", 4), ("relobs2", 4),
                                    ("subjective2", 4), ("pointmass", 2), ("hand", 2), ("delay2", 2)])
def test_step_functions_match_oracle(lib, name, d):
    case = H.Case(name, S=2, T=80, N=5, d=d, weights=True)
    H.check_gains(lib, CPU, case, torch.float64, rtol=1e-9)
    H.check_fwd(lib, CPU, case, torch.float64)
    H.check_vjp(lib, CPU, case, torch.float64)
    H.check_vjp(lib, CPU, case, torch.float32)


def test_large_system_c4(lib):
    """BASELINE config c4 model: TemporalDelayModel(PointMassBoundedActor, delay=2), joint dim 24 (rolled-loop build on the GPU)."""
    case = H.Case("pmdelay2", S=2, T=60, N=4, d=2, weights=True)
    H.check_gains(lib, CPU, case, torch.float64, rtol=1e-9)
    H.check_vjp(lib, CPU, case, torch.float64)
    H.check_vjp(lib, CPU, case, torch.float32)


def test_large_system_c4_full_horizon_parameter_gradients(lib):
    """c4 horizon (T=600, 50 trials): parameter gradients of the FP32 per-trial arithmetic within rtol 1e-3."""
    case = H.Case("pmdelay2", S=2, T=600, N=50, d=2)
    assert H.check_param_vjp(lib, CPU, case, torch.float32) < 1e-3


def test_long_horizon_gradient_tolerance(lib):
    """T=1200 (configs c2/c3 horizon): FP64 per-sample + FP32 per-trial split meets rtol 1e-3 on gradients."""
    case = H.Case("subjective2", S=1, T=1200, N=6, seed=3)
    worst = H.check_vjp(lib, CPU, case, torch.float32)
    assert worst < 1e-3

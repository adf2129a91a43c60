"""Edge cases of the C ABI: time-varying specs (forward), user Sigma0, tiny sizes, error codes.
CPU: host emulation of the step functions; GPU: the CUDA library (marked gpu)."""
import numpy as np
import pytest
import torch

from lqg_b200 import abi
from oracle import adjoint_np as AD
from oracle import lqg_np as O
from tests import helpers as H


def _libs():
    out = [pytest.param("emul", id="emul")]
    out.append(pytest.param("cuda", id="cuda", marks=pytest.mark.gpu))
    return out


def _get(kind):
    if kind == "emul":
        return abi.Library(H.EMUL_PATH), torch.device("cpu")
    return abi.load_library(), torch.device("cuda:0")


@pytest.mark.parametrize("kind", _libs())
def test_time_varying_spec_forward(kind):
    """A genuinely time-varying spec (W, Q, R change with t) through the forward entry points vs the NumPy oracle."""
    lib, dev = _get(kind)
    T, N = 40, 4
    act, dyn = O.make_system(O.bounded_actor_mats(), T)
    rng = np.random.default_rng(0)
    for t in range(T):
        s = 1.0 + 0.5 * np.sin(0.3 * t)
        act["W"][t] = act["W"][t] * s
        act["Q"][t] = act["Q"][t] * (1.0 + 0.02 * t)
        act["R"][t] = act["R"][t] * (1.0 + 0.01 * t)
        dyn["W"][t] = dyn["W"][t] * (2.0 - s)
        dyn["V"][t] = dyn["V"][t] * (1.0 + 0.1 * np.cos(0.2 * t))
    act["Qf"] = act["Q"][-1].copy()
    X = O.simulate(act, dyn, N, rng).astype(np.float32)
    ll_ref = O.log_likelihood(act, dyn, X.astype(np.float64))
    Lo, _, _ = O.lqr_backward(act)
    Ko = O.kf_forward(act, act["V"][0] @ act["V"][0].T)
    dims = abi.LqgkDims(1, N, T, 2, 2, 1, 2, 2)
    ta = {k: torch.tensor(act[k], dtype=torch.float64, device=dev)[None] for k in abi.ACTOR_KEYS}      # (1, T, r, c)
    td = {k: torch.tensor(dyn[k], dtype=torch.float64, device=dev)[None] for k in abi.DYN_KEYS}
    x_tm = lib.pack_obs(torch.tensor(X, device=dev), stream=H.stream_of(dev))
    ws = H.workspace(lib, dims, abi.MODE_FWD, dev)
    wsg = H.workspace(lib, dims, abi.MODE_GAINS, dev)
    if dev.type == "cuda":   # time-varying constants need T constant blocks: give the library generous room
        ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
        wsg = ws
    L, _, _ = lib.lqr_backward(dims, ta, ws=wsg, stream=H.stream_of(dev))
    K = lib.kf_forward(dims, ta, ws=wsg, stream=H.stream_of(dev))
    ll = lib.loglik_fwd(dims, ta, td, x_tm, ws=ws, stream=H.stream_of(dev))
    assert np.allclose(L[0].cpu().numpy(), Lo, rtol=1e-9, atol=1e-12)
    assert np.allclose(K[0].cpu().numpy(), Ko, rtol=1e-9, atol=1e-12)
    assert np.allclose(ll[0].cpu().numpy(), ll_ref, rtol=1e-4)
    with pytest.raises(abi.LqgkError, match="unsupported"):
        lib.loglik_vjp(dims, ta, td, x_tm, ws=ws, stream=H.stream_of(dev))


@pytest.mark.parametrize("kind", _libs())
def test_user_sigma0_value_and_gradient(kind):
    lib, dev = _get(kind)
    T, N = 60, 5
    mats = O.subjective_actor_mats(dim=1)
    sa, sd = O.make_system(mats, T)
    X = O.simulate(sa, sd, N, np.random.default_rng(1)).astype(np.float32)
    S0 = np.array([[2.0, 0.3, 0.1], [0.3, 1.5, 0.2], [0.1, 0.2, 0.8]])
    ll_ref, (ga, gd) = AD.value_and_grad(mats[0], mats[1], X.astype(np.float64), Sigma0=S0)
    assert np.allclose(ll_ref, O.log_likelihood(sa, sd, X.astype(np.float64), Sigma0=S0), rtol=1e-10)
    dims = abi.LqgkDims(1, N, T, 2, 3, 1, 2, 2)
    ta = {k: torch.tensor(np.ascontiguousarray(mats[0][k]), dtype=torch.float64, device=dev)[None] for k in abi.ACTOR_KEYS}
    td = {k: torch.tensor(np.ascontiguousarray(mats[1][k]), dtype=torch.float64, device=dev)[None] for k in abi.DYN_KEYS}
    x_tm = lib.pack_obs(torch.tensor(X, device=dev), stream=H.stream_of(dev))
    ws = H.workspace(lib, dims, abi.MODE_VJP, dev)
    s0 = torch.tensor(S0, dtype=torch.float64, device=dev)[None]
    ll, oa, od, gs0 = lib.loglik_vjp(dims, ta, td, x_tm, sigma0=s0, ws=ws, stream=H.stream_of(dev))
    assert np.allclose(ll[0].cpu().numpy(), ll_ref, rtol=1e-4)
    assert H.rel_err(oa["V"][0].cpu().numpy(), ga["V"]) < 1e-3      # V no longer receives the Sigma0 = V V^T path
    # finite-difference check of the Sigma0 gradient (symmetric perturbation)
    h = 1e-5
    E = np.zeros((3, 3)); E[0, 1] = E[1, 0] = 1.0
    fp = AD.forward(mats[0], mats[1], X.astype(np.float64), Sigma0=S0 + h * E)[0].sum()
    fm = AD.forward(mats[0], mats[1], X.astype(np.float64), Sigma0=S0 - h * E)[0].sum()
    g01 = gs0[0].cpu().numpy()
    assert np.isclose(g01[0, 1] + g01[1, 0], (fp - fm) / (2 * h), rtol=2e-3, atol=1e-6)


@pytest.mark.parametrize("kind", _libs())
def test_tiny_problem_sizes(kind):
    lib, dev = _get(kind)
    for T, N in [(1, 1), (2, 3), (9, 1)]:
        case = H.Case("bounded", S=1, T=T, N=N, seed=T + N)
        H.check_vjp(lib, dev, case, torch.float64)


@pytest.mark.gpu
def test_error_codes_on_gpu():
    lib, dev = abi.load_library(), torch.device("cuda:0")
    case = H.Case("bounded", S=2, T=10, N=3, want_grad=False)
    act, dyn = case.tensors(dev, torch.float32)
    x_tm = lib.pack_obs(torch.tensor(case.X, device=dev))
    bad = abi.LqgkDims(2, 3, 10, 3, 7, 1, 2, 2)
    with pytest.raises(abi.LqgkError, match="unsupported"):
        lib.loglik_fwd(bad, act, dyn, torch.zeros(11, 3, 2, device=dev), ws=torch.empty(1 << 20, dtype=torch.uint8, device=dev))
    dims = case.lqgk_dims()
    with pytest.raises(abi.LqgkError, match="workspace"):
        lib.loglik_fwd(dims, act, dyn, x_tm, ws=torch.empty(256, dtype=torch.uint8, device=dev))
    with pytest.raises(abi.LqgkError, match="invalid"):
        lib.loglik_fwd(dims, act, dyn, x_tm, ws=None)

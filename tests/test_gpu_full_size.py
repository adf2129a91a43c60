"""BASELINE-size checks on the GPU (config c3 shapes: N=100 trials x T=1200) through size-independent properties, plus an
oracle spot check of two samples out of a large batch."""
import numpy as np
import pytest
import torch

import bench
from lqg_b200 import System, abi, runtime, tracking
from oracle import adjoint_np as AD
from oracle import lqg_np as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
N, T = 100, 1200


@pytest.fixture(scope="module")
def data():
    return torch.tensor(bench.make_data(N, T), device=DEV)


def _model(theta):
    return tracking.SubjectiveActor(dim=2, T=T, device=DEV, **{n: theta[:, i] for i, n in enumerate(bench.PARAM_NAMES)})


def _vjp(model_sys, x, S, max_chunk=0, ll_bar=None):
    lib = abi.load_library()
    n, _, d = x.shape
    dims = runtime.dims_of(model_sys.actor, model_sys.dynamics, n, d)
    act = {k: runtime._row_major(getattr(model_sys.actor, k)[:, 0]) for k in abi.ACTOR_KEYS}
    dyn = {k: runtime._row_major(getattr(model_sys.dynamics, k)[:, 0]) for k in abi.DYN_KEYS}
    x_tm = lib.pack_obs(x.contiguous())
    ws = torch.empty(lib.workspace_bytes(dims, abi.MODE_VJP, max_chunk), dtype=torch.uint8, device=DEV)
    out = lib.loglik_vjp(dims, act, dyn, x_tm, ll_bar=ll_bar, ws=ws, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out


def test_chunking_and_trial_permutation_invariance(data):
    S = 2048
    theta = torch.tensor(bench.make_theta(S, 3), device=DEV)
    axis = _model(theta)._axis_system
    xa = data.reshape(N, T + 1, 2, 2).permute(2, 0, 1, 3).reshape(2 * N, T + 1, 2)
    ll0, ga0, gd0, _ = _vjp(axis, xa, S)
    ll1, ga1, gd1, _ = _vjp(axis, xa, S, max_chunk=512)          # 4 chunks: per-sample results must be bit-identical
    assert torch.equal(ll0, ll1)
    assert all(torch.equal(ga0[k], ga1[k]) for k in ga0) and all(torch.equal(gd0[k], gd1[k]) for k in gd0)
    perm = torch.randperm(2 * N, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0))
    ll2, ga2, gd2, _ = _vjp(axis, xa[perm], S)
    assert torch.equal(ll0[:, perm], ll2)                         # trials are independent: exact permutation
    for k in ga0:                                                  # sums over trials in a different order: rounding only
        assert torch.allclose(ga0[k], ga2[k], rtol=2e-4, atol=2e-4 * ga0[k].abs().max().item())


def test_gradient_is_linear_in_the_cotangent(data):
    S = 256
    theta = torch.tensor(bench.make_theta(S, 4), device=DEV)
    axis = _model(theta)._axis_system
    xa = data.reshape(N, T + 1, 2, 2).permute(2, 0, 1, 3).reshape(2 * N, T + 1, 2)
    g = torch.Generator(device=DEV).manual_seed(1)
    w1 = torch.rand(S, 2 * N, device=DEV, generator=g) + 0.5
    w2 = torch.rand(S, 2 * N, device=DEV, generator=g) + 0.5
    _, a1, _, _ = _vjp(axis, xa, S, ll_bar=w1)
    _, a2, _, _ = _vjp(axis, xa, S, ll_bar=w2)
    _, a3, _, _ = _vjp(axis, xa, S, ll_bar=2.0 * w1 - 0.5 * w2)
    for k in a1:
        ref = 2.0 * a1[k] - 0.5 * a2[k]
        assert torch.allclose(a3[k], ref, rtol=1e-3, atol=1e-4 * ref.abs().max().item() + 1e-6), k


def test_factorised_axes_equal_general_kernels_and_oracle(data):
    S = 64
    theta = torch.tensor(bench.make_theta(S, 5), device=DEV, requires_grad=True)
    m = _model(theta)
    ll = m.log_likelihood(data)                                   # per-axis factorisation (n = 5 kernels)
    ll.sum().backward()
    g_fac = theta.grad.clone()
    theta2 = theta.detach().clone().requires_grad_()
    m2 = _model(theta2)
    ll_gen = System(m2.actor, m2.dynamics).log_likelihood(data)   # general n = 10 kernels
    ll_gen.sum().backward()
    assert torch.allclose(ll, ll_gen, rtol=2e-5)
    assert torch.allclose(g_fac, theta2.grad, rtol=1e-3, atol=1e-5 * g_fac.abs().max().item())
    # oracle spot check of two samples (float64, hand adjoint chained to theta through the float64 torch builders)
    from oracle import lqg_torch as OT
    X64 = data.double().cpu().numpy()
    for s in (0, S - 1):
        kw = dict(zip(bench.PARAM_NAMES, theta[s].detach().double().cpu().tolist()))
        mats = O.subjective_actor_mats(dim=2, **kw)
        ll_ref, (ga, gd) = AD.value_and_grad(mats[0], mats[1], X64)
        assert np.allclose(ll[s].detach().cpu().numpy(), ll_ref, rtol=1e-4)
        th = [torch.tensor(kw[k], dtype=torch.float64, requires_grad=True) for k in bench.PARAM_NAMES]
        a, d = OT.subjective_actor(dim=2, **dict(zip(bench.PARAM_NAMES, th)))
        outs = [a[k] for k in "ABFVWQR" if a[k].requires_grad] + [d[k] for k in "ABFVW" if d[k].requires_grad]
        cots = [torch.tensor(ga[k]) for k in "ABFVWQR" if a[k].requires_grad] + [torch.tensor(gd[k]) for k in "ABFVW" if d[k].requires_grad]
        g_ref = torch.autograd.grad(outs, th, grad_outputs=cots, allow_unused=True)
        g_ref = np.array([0.0 if gi is None else gi.item() for gi in g_ref])
        assert np.allclose(g_fac[s].cpu().numpy(), g_ref, rtol=1e-3, atol=1e-5 * np.abs(g_ref).max()), (g_fac[s], g_ref)


def test_32_random_samples_of_a_32768_sample_sweep_match_the_oracle(data):
    """Half of config c3 in ONE call (32,768 parameter samples x 100 trials x T=1200: thread-per-sample covariance kernels,
    7 trials per lane, checkpointed adjoint) through the public API; 32 randomly chosen samples are compared with the float64
    autodiff oracle: log-likelihood rtol 1e-4, every gradient component rtol 1e-3 (components below 1e-4 of the largest one
    on that absolute band)."""
    from oracle import lqg_torch as OT
    S = 32768
    theta = torch.tensor(bench.make_theta(S, 29), device=DEV, requires_grad=True)
    ll = _model(theta).log_likelihood(data)
    ll.sum().backward()
    torch.cuda.synchronize()
    assert ll.shape == (S, N) and torch.isfinite(ll).all() and torch.isfinite(theta.grad).all()
    pick = np.sort(np.random.default_rng(31).choice(S, 32, replace=False))
    tho = theta.detach()[pick].double().cpu().requires_grad_()
    a, d = OT.subjective_actor(dim=2, **{n: tho[:, i] for i, n in enumerate(bench.PARAM_NAMES)})
    llo = OT.log_likelihood(a, d, data.double().cpu())
    llo.sum().backward()
    assert np.allclose(ll.detach()[pick].cpu().numpy(), llo.detach().numpy(), rtol=1e-4)
    g, go = theta.grad[pick].double().cpu().numpy(), tho.grad.numpy()
    band = 1e-3 * np.maximum(np.abs(go), 1e-4 * np.abs(go).max(axis=1, keepdims=True))
    assert (np.abs(g - go) <= band).all(), float((np.abs(g - go) / band).max())

"""GPU tests of the two callers either side of the likelihood that have their own kernels:
lqgk_moments_* (System.conditional_moments / belief_tracking_distribution, lqg/system.py:142-235, 250-257) against the
float64 oracle and the reference-made fixtures, and lqgk_simulate_* (System.simulate, lqg/system.py:62-140) against the
recursion it must satisfy exactly and, in distribution, against the oracle's simulator."""
import glob
import os

import numpy as np
import pytest
import torch

from lqg_b200 import System, tracking
from oracle import lqg_np as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.mark.parametrize("name,cls,kw", [("bounded", tracking.BoundedActor, {}), ("subjective", tracking.SubjectiveActor, {}),
                                         ("subjective2", tracking.SubjectiveActor, {"dim": 2}),
                                         ("relobs", tracking.RelativeObservationBoundedActor, {})])
def test_moments_kernel_matches_oracle(name, cls, kw):
    T, N = 140, 6
    mats = H.model_mats(name)
    sa, sd = O.make_system(mats, T)
    X = O.simulate(sa, sd, N, np.random.default_rng(3))
    m = cls(T=T, device=DEV, dtype=torch.float64, **kw)
    mu, Sig = m._moments(torch.tensor(X, device=DEV))
    n = m.xdim + m.bdim
    assert mu.shape == (N, T, n) and Sig.shape == (T, n, n)
    for i in (0, N - 1):
        mu_o, Sig_o = O.conditional_moments(sa, sd, X[i])
        # per-trial means run in FP32 inside the kernel (like the likelihood); covariances in FP64
        assert np.allclose(mu[i].cpu().numpy(), mu_o, rtol=2e-4, atol=2e-4 * np.abs(mu_o).max())
        assert np.allclose(Sig.cpu().numpy(), Sig_o, rtol=1e-8, atol=1e-9 * np.abs(Sig_o).max())
    # same numbers as the host-side slow path, and the public wrappers
    mu_t, Sig_t = m._moments_torch(torch.tensor(X, device=DEV))
    assert torch.allclose(mu, mu_t.to(mu.dtype), rtol=2e-4, atol=2e-4 * mu_t.abs().max().item())
    assert torch.allclose(Sig, Sig_t.to(Sig.dtype), rtol=1e-8, atol=1e-9 * Sig_t.abs().max().item())
    bt = m.belief_tracking_distribution(torch.tensor(X, device=DEV))
    assert bt.shape() == (N, T, m.bdim)
    mu1, Sig1 = m.conditional_moments(torch.tensor(X[0], device=DEV))
    assert mu1.shape == (T, n) and Sig1.shape == (T, n, n)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(H.ROOT, "tests", "golden", "ref_*.npz"))),
                         ids=lambda p: os.path.basename(p)[4:-4])
def test_moments_kernel_matches_reference_fixtures(path):
    """mu / Sigma of trial 0 at the steps the fixture stores, as computed by the REFERENCE's conditional_moments."""
    from tests.test_reference_golden import Fixture
    fx = Fixture(path)
    m, _ = fx.product_model(DEV, torch.float64)
    if m.xdim + m.bdim > 12:
        pytest.skip("large systems keep the host-side slow path for moments")
    mu, Sig = m._moments(torch.tensor(fx.X, device=DEV))
    st = fx.z["steps"]
    mu_ref, Sig_ref = fx.z["mu_steps"], fx.z["Sigma_steps"]
    assert np.allclose(mu[0].cpu().numpy()[st], mu_ref, rtol=3e-4, atol=3e-4 * np.abs(mu_ref).max())
    assert np.allclose(Sig.cpu().numpy()[st], Sig_ref, rtol=1e-7, atol=1e-8 * np.abs(Sig_ref).max())


def test_batched_moments_one_call():
    T, N, S = 60, 4, 5
    sig = torch.linspace(4.0, 12.0, S, device=DEV, dtype=torch.float64)
    m = tracking.BoundedActor(T=T, sigma_target=sig)
    X = torch.tensor(O.simulate(*O.make_system(O.bounded_actor_mats(), T), N, np.random.default_rng(0)), device=DEV)
    mu, Sig = m._moments(X)
    assert mu.shape == (S, N, T, 4) and Sig.shape == (S, T, 4, 4)
    for s in (0, S - 1):
        sa, sd = O.make_system(O.bounded_actor_mats(sigma_target=float(sig[s])), T)
        mu_o, Sig_o = O.conditional_moments(sa, sd, X[1].cpu().numpy())
        assert np.allclose(Sig[s].cpu().numpy(), Sig_o, rtol=1e-8, atol=1e-9 * np.abs(Sig_o).max())
        assert np.allclose(mu[s, 1].cpu().numpy(), mu_o, rtol=2e-4, atol=2e-4 * np.abs(mu_o).max())


@pytest.mark.parametrize("cls,kw", [(tracking.BoundedActor, {}), (tracking.SubjectiveActor, {"dim": 2}),
                                    (tracking.PointMassBoundedActor, {})])
def test_simulator_satisfies_the_closed_loop_recursion(cls, kw):
    """x, xhat, y, u returned by the kernel obey system.py:106-126 exactly, and the implied noise draws are standard normal."""
    T, n = 200, 4000
    m = cls(T=T, device=DEV, dtype=torch.float64, **kw)
    x, xh, y, u = m.simulate(11, n=n, return_all=True)
    assert x.shape == (n, T + 1, m.xdim) and xh.shape == (n, T + 1, m.bdim) and y.shape == (n, T, m.ydim) and u.shape == (n, T, m.udim)
    gains, K = m._gains()
    a, dn = m.actor, m.dynamics
    mT = lambda M: M.transpose(-1, -2)
    t = T // 3
    assert torch.allclose(u[:, t], xh[:, t] @ mT(gains.L[t]) + gains.l[t], rtol=1e-10, atol=1e-12)
    xp = xh[:, t] @ mT(a.A[t]) + u[:, t] @ mT(a.B[t])
    assert torch.allclose(xh[:, t + 1], xp + (y[:, t] - xp @ mT(a.F[t])) @ mT(K[t]), rtol=1e-9, atol=1e-10)
    # implied standard-normal draws (V, W are invertible in these models): mean 0, identity covariance, white in time
    eps = mT(torch.linalg.solve(dn.V[0], mT(x[:, 1:] - x[:, :-1] @ mT(dn.A[0]) - u @ mT(dn.B[0]))))   # (n, T, x)
    eta = mT(torch.linalg.solve(dn.W[0], mT(y - x[:, 1:] @ mT(dn.F[0]))))                              # (n, T, y)
    for z in (eps, eta):
        flat = z.reshape(-1, z.shape[-1])
        assert flat.mean(0).abs().max() < 5 / np.sqrt(flat.shape[0])
        cov = (mT(flat) @ flat / flat.shape[0]).cpu().numpy()
        assert np.allclose(cov, np.eye(z.shape[-1]), atol=6 / np.sqrt(flat.shape[0]))
        lag1 = (z[:, 1:] * z[:, :-1]).mean().abs().item()
        assert lag1 < 5 / np.sqrt(z[:, 1:].numel())
    # different trials and different seeds give different draws; the same seed reproduces
    assert not torch.equal(x[0], x[1])
    assert torch.equal(m.simulate(11, n=8), x[:8]) and not torch.equal(m.simulate(12, n=8), x[:8])


def test_simulator_matches_the_oracle_in_distribution():
    """First two moments of x_t at t = T/2 and T over 20,000 trials against the oracle's NumPy simulator."""
    T, n = 120, 20000
    m = tracking.SubjectiveActor(T=T, device=DEV, dtype=torch.float64, sigma_target=9.0)
    x = m.simulate(5, n=n).cpu().numpy()
    sa, sd = O.make_system(O.subjective_actor_mats(sigma_target=9.0), T)
    xo = O.simulate(sa, sd, n, np.random.default_rng(9))
    for t in (T // 2, T):
        so = xo[:, t].std(0)
        assert np.all(np.abs(x[:, t].mean(0) - xo[:, t].mean(0)) < 6 * so / np.sqrt(n))
        c, co = np.cov(x[:, t].T), np.cov(xo[:, t].T)
        assert np.allclose(c, co, rtol=0.06, atol=0.06 * np.abs(co).max())


def test_batched_simulation():
    S, T, n = 3, 50, 16
    ac = torch.tensor([0.1, 1.0, 10.0], device=DEV)
    x = tracking.BoundedActor(T=T, action_cost=ac).simulate(3, n=n)
    assert x.shape == (S, n, T + 1, 2) and torch.isfinite(x).all()
    # cheaper control (smaller action cost) tracks the target more tightly
    err = (x[..., 0] - x[..., 1]).pow(2).mean((1, 2))
    assert err[0] < err[1] < err[2]

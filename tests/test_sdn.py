"""Signal-dependent-noise gains (Todorov 2005): an extension with NO reference counterpart (parity unpinned by the reference;
oracle/sdn_np.py restates the published algorithm).  not-gpu: the oracle is pinned by Monte Carlo simulation of the controlled
system and by its reduction to the reference's lqr.backward / kf.forward when the multiplicative noise vanishes.
gpu: the one-kernel CUDA implementation against the oracle."""
import numpy as np
import pytest
import torch

from oracle import lqg_np as O
from oracle import sdn_np as S

T = 60


def test_expected_cost_matches_monte_carlo():
    p = S.tracking_example(c_mult=20.0, d_mult=0.5)
    L, K, cost = S.solve(T=T, sweeps=10, **p)
    m, se = S.simulate_cost(L=L, K=K, n=200000, rng=np.random.default_rng(0), **p)
    assert abs(m - cost) < 4 * se, (m, se, cost)


def test_alternation_converges_and_lowers_cost():
    p = S.tracking_example(c_mult=40.0, d_mult=1.0)
    c1 = S.solve(T=T, sweeps=1, **p)[2]
    L8, K8, c8 = S.solve(T=T, sweeps=8, **p)
    L9, K9, c9 = S.solve(T=T, sweeps=9, **p)
    assert c8 <= c1 + 1e-9 and abs(c9 - c8) < 1e-9 * abs(c8)
    assert np.abs(L9 - L8).max() < 1e-8 and np.abs(K9 - K8).max() < 1e-8


def test_without_multiplicative_noise_reduces_to_reference_gains():
    """C = D = 0: one sweep gives L = -lqr.backward(...).L (lqg/control/lqr.py:16-42; u = -L xhat here) and, with
    Sigma1 = A Sigma0 A' + Om_xi, K_t = A K_t^{kf.forward} (predictor vs filter form, lqg/belief/kf.py:6-21)."""
    p = S.tracking_example()
    V, W = np.linalg.cholesky(p["Om_xi"]), np.linalg.cholesky(p["Om_om"])
    spec = O.time_stack_spec(p["A"], p["B"], p["H"], V, W, p["Q"], p["R"], T)
    Sigma0 = V @ V.T
    p0 = dict(p, C=[], D=[], Sigma1=p["A"] @ Sigma0 @ p["A"].T + p["Om_xi"])
    L, K, _ = S.solve(T=T, sweeps=1, **p0)
    L3, K3, _ = S.solve(T=T, sweeps=3, **p0)
    assert np.abs(L - L3).max() == 0.0 and np.abs(K - K3).max() == 0.0
    Lr, _, _ = O.lqr_backward(spec)
    Kr = O.kf_forward(spec, Sigma0)
    assert np.allclose(L, -Lr, rtol=1e-12, atol=1e-14)
    assert np.allclose(K, p["A"] @ Kr, rtol=1e-10, atol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("sweeps", [0, 1, 6])
def test_cuda_kernel_matches_oracle(sweeps):
    from lqg_b200.control import sdn
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    ps = [S.tracking_example(c_mult=20.0 * np.exp(0.3 * rng.standard_normal()), d_mult=0.5 * np.exp(0.3 * rng.standard_normal()),
                             action_cost=float(np.exp(0.3 * rng.standard_normal())), sigma_target=6.0 * np.exp(0.2 * rng.standard_normal()))
          for _ in range(37)]
    t = lambda k: torch.tensor(np.stack([np.stack(p[k]) if isinstance(p[k], list) else p[k] for p in ps]), device=dev)
    g = sdn.solve(A=t("A")[0], B=t("B")[0], H=t("H")[0], Q=t("Q")[0], R=t("R"), Om_xi=t("Om_xi")[0], Om_omega=t("Om_om"),
                  Sigma1=t("Sigma1")[0], xhat1=t("xhat1")[0], T=T, C=t("C"), D=t("D"), sweeps=sweeps)
    torch.cuda.synchronize()
    for s, p in enumerate(ps):
        L, K, cost = S.solve(T=T, sweeps=sweeps, **p)
        assert np.allclose(g.L[s].cpu().numpy(), L, rtol=1e-9, atol=1e-12)
        if sweeps:
            assert np.allclose(g.K[s].cpu().numpy(), K, rtol=1e-9, atol=1e-12)
        assert np.isclose(g.cost[s].item(), cost, rtol=1e-10)


@pytest.mark.gpu
def test_cuda_kernel_two_dimensional_system():
    """b=4, u=2, y=4 (2-D tracking), no multiplicative noise: equals the CUDA lqr.backward / kf.forward gains."""
    from lqg_b200.control import sdn
    dev = torch.device("cuda:0")
    a, _ = O.bounded_actor_mats(dim=2, action_cost=0.7)
    Om_xi, Om_om = a["V"] @ a["V"].T, a["W"] @ a["W"].T
    t = lambda M: torch.tensor(np.ascontiguousarray(M), device=dev)
    Sigma1 = a["A"] @ Om_xi @ a["A"].T + Om_xi
    g = sdn.solve(A=t(a["A"]), B=t(a["B"]), H=t(a["F"]), Q=t(a["Q"]), R=t(a["R"]), Om_xi=t(Om_xi), Om_omega=t(Om_om), Sigma1=t(Sigma1),
                  xhat1=torch.zeros(4, dtype=torch.float64, device=dev), T=T, sweeps=1)
    spec = O.time_stack_spec(a["A"], a["B"], a["F"], a["V"], a["W"], a["Q"], a["R"], T)
    Lr, _, _ = O.lqr_backward(spec)
    Kr = O.kf_forward(spec, Om_xi)
    assert np.allclose(g.L[0].cpu().numpy(), -Lr, rtol=1e-9, atol=1e-12)
    assert np.allclose(g.K[0].cpu().numpy(), a["A"] @ Kr, rtol=1e-8, atol=1e-11)

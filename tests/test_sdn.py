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


# ------------------------------------------------------------------------------------------------ filter form (the reference's conventions)
def _filter_kw(p):
    return dict(A=p["A"], B=p["B"], F=p["H"], Q=p["Q"], R=p["R"], Qf=p["Qf"], Om_xi=p["Om_xi"], Om_om=p["Om_om"])


def test_filter_form_without_multiplicative_noise_is_exactly_the_reference():
    """C = D = 0: ONE sweep gives lqr.backward's L (same sign, lqg/control/lqr.py:16-42) and kf.forward's K (lqg/belief/kf.py:6-21)."""
    p = S.tracking_example()
    V, W = np.linalg.cholesky(p["Om_xi"]), np.linalg.cholesky(p["Om_om"])
    spec = O.time_stack_spec(p["A"], p["B"], p["H"], V, W, p["Q"], p["R"], T)
    Sigma0 = V @ V.T
    L, K, _ = S.filter_solve(C=[], D=[], Sigma0=Sigma0, xhat0=p["xhat1"], T=T, sweeps=1, **_filter_kw(p))
    Lr, _, _ = O.lqr_backward(spec)
    Kr = O.kf_forward(spec, Sigma0)
    assert np.allclose(L, Lr, rtol=1e-12, atol=1e-14) and np.allclose(K, Kr, rtol=1e-12, atol=1e-14)


def test_filter_form_cost_monte_carlo_exact_evaluation_and_optimality():
    p = S.tracking_example(c_mult=20.0, d_mult=0.5)
    kw = dict(C=p["C"], D=p["D"], Sigma0=p["Sigma1"], xhat0=p["xhat1"], **_filter_kw(p))
    c1 = S.filter_solve(T=T, sweeps=1, **kw)[2]
    L8, K8, c8 = S.filter_solve(T=T, sweeps=8, **kw)
    L9, K9, c9 = S.filter_solve(T=T, sweeps=9, **kw)
    assert c8 <= c1 + 1e-9 and abs(c9 - c8) < 1e-9 * abs(c8) and np.abs(L9 - L8).max() < 1e-8 and np.abs(K9 - K8).max() < 1e-8
    # the value-function cost equals the exact second-moment evaluation of the same gains ...
    ce = S.filter_expected_cost(L=L9, K=K9, **kw)
    assert abs(ce - c9) < 1e-9 * abs(c9)
    # ... and the Monte Carlo cost of the simulated closed loop
    m, se = S.filter_simulate_cost(L=L9, K=K9, n=200000, rng=np.random.default_rng(0), **kw)
    assert abs(m - c9) < 4 * se, (m, se, c9)
    # coordinate-wise optimality: perturbing any single L_t or K_t raises the exactly evaluated cost
    rng = np.random.default_rng(1)
    for t in (0, 7, 31, T - 1):
        L2, K2 = L9.copy(), K9.copy()
        L2[t] += 1e-2 * np.abs(L9[t]).max() * rng.standard_normal(L9[t].shape)
        K2[t] += 1e-2 * max(np.abs(K9[t]).max(), 1e-3) * rng.standard_normal(K9[t].shape)
        assert S.filter_expected_cost(L=L2, K=K9, **kw) > ce and S.filter_expected_cost(L=L9, K=K2, **kw) >= ce
    # knowing about the multiplicative noise pays: the plain LQG gains cost more under it
    Lp, Kp, _ = S.filter_solve(T=T, sweeps=1, **dict(kw, C=[], D=[]))
    assert S.filter_expected_cost(L=Lp, K=Kp, **kw) > 1.05 * ce


def _filter_library_vs_oracle(lib, dev, sweeps, n=9):
    rng = np.random.default_rng(3)
    ps = [S.tracking_example(c_mult=20.0 * np.exp(0.3 * rng.standard_normal()), d_mult=0.5 * np.exp(0.3 * rng.standard_normal()),
                             action_cost=float(np.exp(0.3 * rng.standard_normal())), sigma_target=6.0 * np.exp(0.2 * rng.standard_normal()))
          for _ in range(n)]
    t = lambda k: torch.tensor(np.stack([np.stack(p[k]) if isinstance(p[k], list) else p[k] for p in ps]), device=dev)
    mats = dict(A=t("A")[0], B=t("B")[0], H=t("H")[0], Q=t("Q")[0], R=t("R"), Qf=None, Om_xi=t("Om_xi")[0], Om_omega=t("Om_om"),
                Sigma1=t("Sigma1")[0], xhat1=t("xhat1")[0], C=t("C"), D=t("D"))
    L, K, cost = lib.sdn_gains(mats, T, sweeps, filter_form=True)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    for i, p in enumerate(ps):
        Lr, Kr, cr = S.filter_solve(C=p["C"], D=p["D"], Sigma0=p["Sigma1"], xhat0=p["xhat1"], T=T, sweeps=sweeps, **_filter_kw(p))
        assert np.allclose(L[i].cpu().numpy(), Lr, rtol=1e-9, atol=1e-11), (i, np.abs(L[i].cpu().numpy() - Lr).max())
        assert np.allclose(K[i].cpu().numpy(), Kr, rtol=1e-9, atol=1e-11)
        assert abs(cost[i].item() - cr) < 1e-9 * abs(cr)


@pytest.mark.parametrize("sweeps", [0, 1, 5])
def test_filter_form_step_functions_match_oracle_on_host(sweeps):
    from lqg_b200 import abi
    from tests import helpers as H
    _filter_library_vs_oracle(abi.Library(H.EMUL_PATH), torch.device("cpu"), sweeps)


@pytest.mark.gpu
@pytest.mark.parametrize("sweeps", [0, 1, 6])
def test_filter_form_cuda_kernel_matches_oracle(sweeps):
    from lqg_b200 import abi
    _filter_library_vs_oracle(abi.load_library(), torch.device("cuda:0"), sweeps, n=37)


@pytest.mark.gpu
def test_solve_for_actor_reproduces_reference_gains_and_feeds_the_sdn_likelihood():
    from lqg_b200.control import sdn
    from lqg_b200.tracking import SubjectiveActor
    dev = torch.device("cuda:0")
    model = SubjectiveActor(dim=1, T=80, dtype=torch.float64, device=dev)
    g0 = sdn.solve_for_actor(model, sweeps=1)                                  # no multiplicative noise
    gains, K = model._gains()
    assert torch.allclose(g0.L[0], gains.L, rtol=1e-9, atol=1e-12) and torch.allclose(g0.K[0], K, rtol=1e-9, atol=1e-12)
    g = sdn.solve_for_actor(model, signal_dep_noise=40.0, obs_dep_noise=0.4)
    x = model.simulate(3, n=5)[..., :2].to(torch.float32)
    ll_plain = model.log_likelihood_sdn(x, signal_dep_noise=40.0, obs_dep_noise=0.4)
    ll_sdn = model.log_likelihood_sdn(x, signal_dep_noise=40.0, obs_dep_noise=0.4, gains=(g.L[0], g.K[0]))
    assert torch.isfinite(ll_sdn).all() and not torch.allclose(ll_sdn, ll_plain)

"""GPU tests of the Python drop-in layer (lqg_b200.spec/system/tracking/control/belief): shapes and calls mirror the
reference's tests/lqg_test.py and tests/infer_test.py; values are checked against the float64 oracle."""
import numpy as np
import pytest
import torch

from lqg_b200 import LQG, System, tracking
from lqg_b200.belief import kf
from lqg_b200.control import lqr
from oracle import lqg_np as O
from oracle import lqg_torch as OT

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _sim(mats, T, N, seed=0, d=None):
    sa, sd = O.make_system(mats, T)
    X = O.simulate(sa, sd, N, np.random.default_rng(seed))
    return X if d is None else X[..., :d]


def test_lqg_simulate_shapes():
    """reference tests/lqg_test.py:16-43"""
    dt, T = 1.0 / 60.0, 300
    A = torch.eye(2); B = torch.tensor([[0.0], [dt]]); V = torch.diag(torch.tensor([1.0, 0.5]))
    F = torch.eye(2); W = torch.diag(torch.tensor([6.0, 3.0])); Q = torch.tensor([[1.0, -1.0], [-1.0, 1.0]])
    R = torch.eye(1) * 0.5
    lqg = LQG(A=A.to(DEV), B=B.to(DEV), F=F.to(DEV), V=V.to(DEV), W=W.to(DEV), Q=Q.to(DEV), R=R.to(DEV), T=T)
    x = lqg.simulate(0, x0=torch.zeros(2), n=10)
    assert x.shape == (10, T + 1, 2) and torch.isfinite(x).all()


@pytest.mark.parametrize("cls", [tracking.BoundedActor, tracking.SubjectiveActor, tracking.PointMassBoundedActor,
                                 tracking.OptimalActor, tracking.RelativeObservationBoundedActor])
def test_model_simulates(cls):
    """reference tests/lqg_test.py:46-66"""
    T = 200
    m = cls(T=T, device=DEV)
    x = m.simulate(0, x0=torch.zeros(m.xdim), n=10)
    assert x.shape == (10, T + 1, m.xdim) and not torch.isnan(x).any()


def test_subjective_without_subjective_component_equals_bounded():
    """reference tests/lqg_test.py:69-93 (same generator seed -> identical trajectories)"""
    kw = dict(process_noise=1.0, sigma_target=6.0, action_cost=0.1, action_variability=0.5, sigma_cursor=3.0, T=300,
              device=DEV, dtype=torch.float64)
    xb = tracking.BoundedActor(**kw).simulate(0, n=20)
    xs = tracking.SubjectiveActor(subj_noise=1.0, subj_vel_noise=0.0, **kw).simulate(0, n=20)
    assert torch.allclose(xb, xs, rtol=1e-8, atol=1e-8)


def test_conditional_distribution_and_belief_shapes():
    """reference tests/infer_test.py:10-16 and tests/lqg_test.py:96-106"""
    T = 120
    m = tracking.SubjectiveActor(T=T, device=DEV)
    x = m.simulate(113, n=20)
    cd = m.conditional_distribution(x)
    assert cd.shape()[1] == x.shape[1] - 1
    lp = cd.log_prob(x[:, 1:])
    assert lp.shape == (20,) and torch.isfinite(lp).all()
    a = tracking.BoundedActor(T=T, device=DEV)
    xa = a.simulate(0, n=20)
    assert a.belief_tracking_distribution(xa).shape() == (20, T, a.actor.A.shape[-1])
    mu, Sig = a.conditional_moments(xa[0])
    assert mu.shape == (T, 4) and Sig.shape == (T, 4, 4)


def test_numpyro_adaptor_contract():
    """reference tests/infer_test.py:29-47 (distribution part)"""
    T = 100
    m = tracking.BoundedActor(T=T, device=DEV)
    dist = m.to_numpyro()
    assert dist.event_shape == (T + 1, 2) and dist.batch_shape == ()
    x = dist.sample(0, sample_shape=(10,))
    assert x.shape == (10, T + 1, 2)
    assert dist.sample(2).shape == (T + 1, 2)
    assert torch.isfinite(dist.log_prob(x)).all()


def test_gains_api_matches_oracle():
    T = 250
    m = tracking.SubjectiveActor(dim=2, T=T, device=DEV, dtype=torch.float64, action_cost=0.3)
    g = lqr.backward(m.actor)
    K = kf.forward(m.actor, m.actor.V[0] @ m.actor.V[0].T)
    sa, _ = O.make_system(O.subjective_actor_mats(dim=2, action_cost=0.3), T)
    Lo, lo, Ho = O.lqr_backward(sa)
    Ko = O.kf_forward(sa, sa["V"][0] @ sa["V"][0].T)
    assert g.L.shape == (T, 2, 6) and g.l.shape == (T, 2) and g.H.shape == (T, 2, 2) and K.shape == (T, 6, 4)
    assert np.allclose(g.L.cpu().numpy(), Lo, rtol=1e-9, atol=1e-12) and np.allclose(K.cpu().numpy(), Ko, rtol=1e-9, atol=1e-12)
    assert np.allclose(g.H.cpu().numpy(), Ho, rtol=1e-9)


@pytest.mark.parametrize("dim", [1, 2])
def test_log_likelihood_and_parameter_gradient(dim):
    """jax.value_and_grad(lambda th: Model(**th).log_likelihood(x).sum()) of the reference == autograd through the
    torch model builders + CUDA adjoint here; checked against the float64 autodiff oracle (rtol 1e-4 / 1e-3)."""
    T, N, S = 150, 12, 5
    X = _sim(O.subjective_actor_mats(dim=dim, sigma_target=9.0), T, N, seed=4).astype(np.float32)
    names = ["action_cost", "action_variability", "subj_noise", "subj_vel_noise", "sigma_target", "sigma_cursor"]
    th0 = np.array([1.0, 0.5, 1.0, 0.5, 9.0, 6.0])
    th_np = th0[None] * np.exp(0.2 * np.random.default_rng(5).standard_normal((S, 6)))
    tho = torch.tensor(th_np, dtype=torch.float64, requires_grad=True)
    a, d = OT.subjective_actor(dim=dim, **{n: tho[:, i] for i, n in enumerate(names)})
    llo = OT.log_likelihood(a, d, torch.tensor(X, dtype=torch.float64))
    llo.sum().backward()
    th = torch.tensor(th_np, dtype=torch.float32, device=DEV, requires_grad=True)
    m = tracking.SubjectiveActor(dim=dim, T=T, **{n: th[:, i] for i, n in enumerate(names)})
    ll = m.log_likelihood(torch.tensor(X, device=DEV))
    assert ll.shape == (S, N)
    ll.sum().backward()
    assert np.allclose(ll.detach().cpu().numpy(), llo.detach().numpy(), rtol=1e-4)
    g, go = th.grad.cpu().numpy().astype(np.float64), tho.grad.numpy()
    assert np.allclose(g, go, rtol=1e-3, atol=1e-3 * np.abs(go).max(axis=1, keepdims=True) * 1e-2), np.abs(g / go - 1).max()
    if dim == 2:   # the general (un-factorised, n = 10) kernels give the same numbers
        gen = System(m.actor, m.dynamics)
        ll2 = gen.log_likelihood(torch.tensor(X, device=DEV))
        assert np.allclose(ll2.detach().cpu().numpy(), llo.detach().numpy(), rtol=1e-4)


def test_per_trial_weights_and_shared_axes_in_backward():
    """ll_bar that varies over trials takes the explicit-adjoint path; an un-batched model returns (n,)."""
    T, N = 80, 9
    X = torch.tensor(_sim(O.bounded_actor_mats(), T, N, seed=2).astype(np.float32), device=DEV)
    sig = torch.tensor(6.0, device=DEV, requires_grad=True)
    m = tracking.BoundedActor(T=T, sigma_target=sig)
    ll = m.log_likelihood(X)
    assert ll.shape == (N,)
    w = torch.linspace(0.5, 1.5, N, device=DEV)
    (ll * w).sum().backward()
    so = torch.tensor(6.0, dtype=torch.float64, requires_grad=True)
    a, d = OT.bounded_actor(sigma_target=so)
    (OT.log_likelihood(a, d, X.double().cpu()) * w.double().cpu()).sum().backward()
    assert np.isclose(sig.grad.item(), so.grad.item(), rtol=1e-3)


def test_no_cpu_fallback():
    m = tracking.BoundedActor(T=10, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.log_likelihood(torch.zeros(2, 11, 2))


def test_per_condition_data_in_one_call():
    """Config c2 shape: one sample per blob-width condition, each with its own trials (reference loops over
    conditions, lqg/infer/models.py:38-61; here the conditions are the kernels' sample axis)."""
    T, N, C = 120, 7, 3
    sig = [8.5, 19.9, 51.6]
    Xs, lls, grads = [], [], []
    for c in range(C):
        mats = O.subjective_actor_mats(dim=2, sigma_target=sig[c])
        X = _sim(mats, T, N, seed=10 + c).astype(np.float32)
        Xs.append(X)
        so = torch.tensor(sig[c], dtype=torch.float64, requires_grad=True)
        a, d = OT.subjective_actor(dim=2, sigma_target=so)
        ll = OT.log_likelihood(a, d, torch.tensor(X, dtype=torch.float64))
        ll.sum().backward()
        lls.append(ll.detach().numpy()); grads.append(so.grad.item())
    st = torch.tensor(sig, device=DEV, requires_grad=True)
    m = tracking.SubjectiveActor(dim=2, T=T, sigma_target=st)
    x = torch.tensor(np.stack(Xs), device=DEV)                       # (C, N, T+1, 4)
    ll = m.log_likelihood(x)
    assert ll.shape == (C, N)
    ll.sum().backward()
    assert np.allclose(ll.detach().cpu().numpy(), np.stack(lls), rtol=1e-4)
    assert np.allclose(st.grad.cpu().numpy(), np.array(grads), rtol=1e-3)


def test_per_condition_data_odd_stride_dim1():
    """d = 2 with an odd number of (trial, step) rows per condition: the per-sample stride n*(T+1)*d floats is only 8-byte
    aligned, which is all the float2 observation accesses need (was rejected as 'invalid argument')."""
    T, N, C = 120, 7, 3                                               # N * (T + 1) = 847 is odd
    sig = [5.0, 9.0, 14.0]
    Xs, lls = [], []
    for c in range(C):
        mats = O.bounded_actor_mats(sigma_target=sig[c])
        X = _sim(mats, T, N, seed=20 + c).astype(np.float32)
        Xs.append(X)
        sa, sd = O.make_system(mats, T)
        lls.append(O.log_likelihood(sa, sd, X.astype(np.float64)))
    m = tracking.BoundedActor(dim=1, T=T, sigma_target=torch.tensor(sig, device=DEV))
    ll = m.log_likelihood(torch.tensor(np.stack(Xs), device=DEV))     # (C, N, T+1, 2)
    assert ll.shape == (C, N)
    assert np.allclose(ll.cpu().numpy(), np.stack(lls), rtol=1e-4)


def test_to_dtype_moves_the_factorised_axis_model():
    """System.to() must convert the cached 1-axis model the dim > 1 likelihood runs on."""
    T, N = 60, 5
    X = torch.tensor(_sim(O.bounded_actor_mats(dim=2), T, N, seed=3), device=DEV)
    m32 = tracking.BoundedActor(dim=2, T=T, device=DEV)
    m64 = m32.to(torch.float64)
    assert m64._axis_system.dtype == torch.float64 and m64.dtype == torch.float64
    ll64 = m64.log_likelihood(X)
    assert ll64.dtype == torch.float64
    gen = System(m64.actor, m64.dynamics).log_likelihood(X)            # un-factorised n = 8 kernels
    assert torch.allclose(ll64, gen, rtol=1e-6)


def test_terminal_cost_is_forwarded_and_cross_cost_refused():
    """a spec whose Qf differs from Q[-1] must reach the kernels (lqr.py:37 starts from spec.Qf); P != 0 is
    outside the fused path and must raise instead of being dropped."""
    T, N = 90, 6
    mats = O.bounded_actor_mats()
    X = _sim(mats, T, N, seed=8).astype(np.float32)
    sa, sd = O.make_system(mats, T)
    Qf = np.array([[3.0, -2.0], [-2.0, 5.0]])
    sa2 = dict(sa, Qf=Qf)
    ll_ref = O.log_likelihood(sa2, sd, X.astype(np.float64))
    assert not np.allclose(ll_ref, O.log_likelihood(sa, sd, X.astype(np.float64)), rtol=1e-6)
    m = tracking.BoundedActor(T=T, device=DEV, dtype=torch.float64)
    Qf_t = torch.tensor(Qf, device=DEV, requires_grad=True)
    s2 = System(m.actor._replace(Qf=Qf_t), m.dynamics)
    ll = s2.log_likelihood(torch.tensor(X, device=DEV))
    assert np.allclose(ll.detach().cpu().numpy(), ll_ref, rtol=1e-6)
    ll.sum().backward()
    assert Qf_t.grad is not None and torch.isfinite(Qf_t.grad).all() and Qf_t.grad.abs().max() > 0
    eps = 1e-5                                                         # finite difference of the oracle in Qf[0, 0]
    llp = O.log_likelihood(dict(sa, Qf=Qf + np.array([[eps, 0], [0, 0]])), sd, X.astype(np.float64)).sum()
    llm = O.log_likelihood(dict(sa, Qf=Qf - np.array([[eps, 0], [0, 0]])), sd, X.astype(np.float64)).sum()
    assert np.isclose(Qf_t.grad[0, 0].item(), (llp - llm) / (2 * eps), rtol=2e-3)
    P = torch.ones_like(m.actor.P)
    with pytest.raises(NotImplementedError, match="P = 0"):
        System(m.actor._replace(P=P), m.dynamics).log_likelihood(torch.tensor(X, device=DEV))


def test_vjp_call_is_capturable_into_a_cuda_graph_and_replays_identically():
    """After lqgk_init() no entry point creates CUDA objects, so one fused forward+adjoint call (with its internal fork/join
    over auxiliary streams) can be captured into a CUDA graph -- how XLA may run an FFI handler -- and replayed."""
    from lqg_b200 import abi
    lib = abi.load_library()
    lib.init(1)
    T, N, S = 100, 10, 40
    mats = [O.subjective_actor_mats(sigma_target=6.0 + s) for s in range(S)]
    X = _sim(mats[0], T, N, seed=6).astype(np.float32)
    dims = abi.LqgkDims(S, N, T, 2, 3, 1, 2, 2)
    act = {k: torch.tensor(np.stack([np.ascontiguousarray(m[0][k]) for m in mats]), dtype=torch.float32, device=DEV) for k in abi.ACTOR_KEYS}
    dyn = {k: torch.tensor(np.stack([np.ascontiguousarray(m[1][k]) for m in mats]), dtype=torch.float32, device=DEV) for k in abi.DYN_KEYS}
    x_tm = lib.pack_obs(torch.tensor(X, device=DEV))
    ws = torch.empty(lib.workspace_bytes(dims, abi.MODE_VJP, 0), dtype=torch.uint8, device=DEV)
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        ll0, ga0, gd0, _ = lib.loglik_vjp(dims, act, dyn, x_tm, ws=ws, stream=st.cuda_stream)   # eager (also sets kernel attributes)
    st.synchronize()
    ref = [ll0.clone()] + [v.clone() for v in list(ga0.values()) + list(gd0.values())]
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=st):
        ll1, ga1, gd1, _ = lib.loglik_vjp(dims, act, dyn, x_tm, ws=ws, stream=torch.cuda.current_stream().cuda_stream)
    outs = [ll1] + list(ga1.values()) + list(gd1.values())
    for rep in range(2):
        for o in outs:
            o.fill_(float("nan"))
        ws.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(outs[0], ref[0])                                     # log-likelihoods: bit-identical
        for o, r in zip(outs[1:], ref[1:]):                                     # gradients: FP64 atomics order may differ in the last bit
            assert torch.allclose(o, r, rtol=1e-6, atol=1e-7 * r.abs().max().item())


def test_cuda_graph_replay_of_value_and_grad_through_the_public_api():
    """Model construction + fused likelihood + backward captured once (lqg_b200.graphs) and replayed with new parameters."""
    from lqg_b200.graphs import GraphedValueAndGrad
    T, N, S = 90, 8, 5
    X = torch.tensor(_sim(O.subjective_actor_mats(dim=2), T, N, seed=12).astype(np.float32), device=DEV)
    names = ["action_cost", "action_variability", "subj_noise", "subj_vel_noise", "sigma_target", "sigma_cursor"]

    def fn(th):
        m = tracking.SubjectiveActor(dim=2, T=T, **{n: th[:, i] for i, n in enumerate(names)})
        return m.log_likelihood(X).sum(-1)

    th0 = torch.tensor([[1.0, 0.5, 1.0, 0.5, 6.0, 6.0]], device=DEV).repeat(S, 1)
    gv = GraphedValueAndGrad(fn, th0)
    for seed in (1, 2):
        th = th0 * torch.exp(0.2 * torch.randn(S, 6, device=DEV, generator=torch.Generator(device=DEV).manual_seed(seed)))
        ll_g, g_g = gv(th)
        ll_g, g_g = ll_g.clone(), g_g.clone()
        th_e = th.clone().requires_grad_()
        ll_e = fn(th_e)
        ll_e.sum().backward()
        assert torch.allclose(ll_g, ll_e.detach(), rtol=1e-6)
        assert torch.allclose(g_g, th_e.grad, rtol=1e-5, atol=1e-6 * th_e.grad.abs().max().item())

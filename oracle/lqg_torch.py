"""Float64 torch restatement, batched over parameter samples, differentiable (TEST ORACLE ONLY).

Same recursions as ``oracle/lqg_np.py`` (reference lqg/control/lqr.py:16-42, lqg/belief/kf.py:6-21,
lqg/system.py:142-248) for *time-invariant* base matrices carrying arbitrary leading batch dims
``[..., r, c]``.  Used (a) as the gradient oracle (torch autograd == what JAX autodiff computes in
the reference), (b) as the CPU baseline / ``bench.py --impl reference`` port (the real
JAX reference is not installable in this image).  Parity unpinned by reference goldens; see
``oracle/__init__.py``.
"""
from __future__ import annotations

import math

import torch

LOG2PI = math.log(2.0 * math.pi)
mT = lambda M: M.transpose(-1, -2)


def _eye_like(n, ref):
    return torch.eye(n, dtype=ref.dtype, device=ref.device)


def lqr_backward(A, B, Q, R, T, eps=1e-8):
    """lqr.py:16-42 with q = r = P = 0 (true for every model, lqg/utils.py:29-33).  L[T,...,u,b]."""
    S = Q
    u = R.shape[-1]
    Ls = []
    for _ in range(T):
        H = R + mT(B) @ S @ B
        G = mT(B) @ S @ A
        lam = torch.linalg.eigvalsh(0.5 * (H + mT(H)))[..., 0]
        Ht = H + torch.clamp(eps - lam, min=0.0)[..., None, None] * _eye_like(u, H)
        L = -torch.linalg.solve(Ht, G)
        S = Q + mT(A) @ S @ A + mT(L) @ H @ L + mT(L) @ G + mT(G) @ L
        Ls.append(L)
    return torch.stack(Ls[::-1])


def kf_forward(A, F, V, W, Sigma0, T):
    """kf.py:6-21.  K[T,...,b,y]."""
    P = Sigma0
    b = A.shape[-1]
    VV, WW = V @ mT(V), W @ mT(W)
    Ks = []
    for _ in range(T):
        P = A @ P @ mT(A) + VV
        G = F @ P @ mT(F) + WW
        K = P @ mT(F) @ torch.linalg.inv(G)
        P = (_eye_like(b, P) - K @ F) @ P
        Ks.append(K)
    return torch.stack(Ks)


def log_likelihood(act, dyn, X, T=None, Sigma0=None):
    """system.py:142-248.  ``act``/``dyn``: dicts of base matrices ``[..., r, c]`` (batch dims broadcast),
    ``X[N, T+1, d]``.  Returns ``ll[..., N]`` (sum over time, per trial)."""
    X = X.to(torch.float64)
    N, T1, d = X.shape
    T = T1 - 1 if T is None else T
    Aa, Ba, Fa, Va, Wa, Q, R = (act[k] for k in ("A", "B", "F", "V", "W", "Q", "R"))
    Ad, Bd, Fd, Vd, Wd = (dyn[k] for k in ("A", "B", "F", "V", "W"))
    x, b, y = Ad.shape[-1], Aa.shape[-1], Wd.shape[-1]
    L = lqr_backward(Aa, Ba, Q, R, T)
    K = kf_forward(Aa, Fa, Va, Wa, Va @ mT(Va) if Sigma0 is None else Sigma0, T)
    batch = torch.broadcast_shapes(L.shape[1:-2], K.shape[1:-2], Ad.shape[:-2], Vd.shape[:-2], Wd.shape[:-2])
    ex = lambda M: M.expand(batch + M.shape[-2:])
    D = Fd @ Bd - Fa @ Ba
    FAd, FAa, FVd = Fd @ Ad, Fa @ Aa, Fd @ Vd

    def joint(t):                                                     # system.py:167-207
        Lt, Kt = L[t], K[t]
        Fj = torch.cat([torch.cat([ex(Ad), ex(Bd @ Lt)], -1),
                        torch.cat([ex(Kt @ FAd), ex(Aa + Ba @ Lt - Kt @ FAa + Kt @ D @ Lt)], -1)], -2)
        Gj = torch.cat([torch.cat([ex(Vd), torch.zeros(batch + (x, y), dtype=X.dtype)], -1),
                        torch.cat([ex(Kt @ FVd), ex(Kt @ Wd)], -1)], -2)
        return Fj, Gj

    n = x + b
    F0, G0 = joint(0)
    Sig = G0 @ mT(G0)                                                  # system.py:212
    mu = torch.zeros(batch + (N, n), dtype=X.dtype)
    mu[..., :d] = X[:, 0]                                              # system.py:211
    ll = torch.zeros(batch + (N,), dtype=X.dtype)
    for t in range(T):
        Fj, Gj = joint(t)
        FS = Fj @ Sig
        Soo = Sig[..., :d, :d]
        r = X[:, t] - mu[..., :d]                                      # [..., N, d]
        w = torch.linalg.solve(Soo, mT(r))                             # [..., d, N]
        mu = mu @ mT(Fj) + mT(FS[..., :, :d] @ w)                      # system.py:219-221
        Sig = FS @ mT(Fj) + Gj @ mT(Gj) - FS[..., :, :d] @ torch.linalg.solve(Soo, (Sig @ mT(Fj))[..., :d, :])  # :223-230
        Lc = torch.linalg.cholesky(Sig[..., :d, :d])                   # numpyro MVN.log_prob
        e = X[:, t + 1] - mu[..., :d]
        z = torch.linalg.solve_triangular(Lc, mT(e), upper=False)      # [..., d, N]
        ll = ll - 0.5 * d * LOG2PI - torch.log(torch.diagonal(Lc, dim1=-2, dim2=-1)).sum(-1)[..., None] \
            - 0.5 * (z * z).sum(-2)
    return ll


# --------------------------------------------------------------------------- differentiable model builders
def _t(v, ref=None):
    return v if torch.is_tensor(v) else torch.tensor(float(v), dtype=torch.float64)


def _diag_from(vals):
    """vals: list of tensors with broadcastable batch shape -> [..., k, k] diagonal matrix."""
    vals = torch.broadcast_tensors(*[_t(v).to(torch.float64) for v in vals])
    return torch.diag_embed(torch.stack(vals, -1))


def _const(M):
    return torch.as_tensor(M, dtype=torch.float64)


def bounded_actor(dim=1, process_noise=1.0, action_variability=0.5, sigma_target=6.0, sigma_cursor=6.0,
                  action_cost=1.0, dt=1.0 / 60.0):
    """tracking/basic.py:7-64, parameters may be tensors with a leading batch shape."""
    from . import lqg_np
    base, _ = lqg_np.bounded_actor_mats(dim=dim, dt=dt)
    act = dict(A=_const(base["A"]), B=_const(base["B"]), F=_const(base["F"]), Q=_const(base["Q"]))
    act["V"] = _diag_from([process_noise, action_variability] * dim)
    act["W"] = _diag_from([sigma_target, sigma_cursor] * dim)
    act["R"] = _t(action_cost).to(torch.float64)[..., None, None] * torch.eye(dim, dtype=torch.float64)
    return act, dict(act)


def subjective_actor(dim=1, process_noise=1.0, action_cost=1.0, action_variability=0.5, subj_noise=1.0,
                     subj_vel_noise=0.5, sigma_target=6.0, sigma_cursor=6.0, dt=1.0 / 60.0):
    """tracking/subjective.py:15-47, parameters may be tensors with a leading batch shape."""
    from . import lqg_np
    a0, d0 = lqg_np.subjective_actor_mats(dim=dim, dt=dt)
    p = lqg_np.swap_dims(3 * dim, dim)
    W = _diag_from([sigma_target, sigma_cursor] * dim)
    dyn = dict(A=_const(d0["A"]), B=_const(d0["B"]), F=_const(d0["F"]), W=W,
               V=_diag_from([process_noise, action_variability] * dim))
    Vfull = _diag_from([subj_noise, action_variability, subj_vel_noise] * dim)
    act = dict(A=_const(a0["A"]), B=_const(a0["B"]), F=_const(a0["F"]), Q=_const(a0["Q"]), W=W,
               V=Vfull[..., p, :],
               R=_t(action_cost).to(torch.float64)[..., None, None] * torch.eye(dim, dtype=torch.float64))
    return act, dyn


# --------------------------------------------------------------------------- signal-dependent noise (extension, oracle/sdn_np.py)
def sdn_log_likelihood(act, dyn, X, c_scale, d_scale, T=None):
    """Differentiable twin of oracle/sdn_np.py: sdn_log_likelihood for ONE (un-batched) parameter vector with the per-channel
    noise model of lqg_b200.control.sdn.channel_noise (C_i = c_scale B_d[:, i] e_i^T, D_j = d_scale e_j e_j^T F_d) and the actor's
    own lqr.backward / kf.forward gains.  The autograd gradient of this function is the independent check of the batched
    central differences the product uses for the signal-dependent-noise likelihood.  X[N, T+1, d] -> ll[N]."""
    X = X.to(torch.float64)
    N, T1, d = X.shape
    T = T1 - 1 if T is None else T
    Aa, Ba, Fa, Va, Wa, Q, R = (act[k] for k in ("A", "B", "F", "V", "W", "Q", "R"))
    Ad, Bd, Fd, Vd, Wd = (dyn[k] for k in ("A", "B", "F", "V", "W"))
    x, b, u, y = Ad.shape[-1], Aa.shape[-1], Bd.shape[-1], Wd.shape[-1]
    L = lqr_backward(Aa, Ba, Q, R, T)
    K = kf_forward(Aa, Fa, Va, Wa, Va @ mT(Va), T)
    eye_u, eye_y = torch.eye(u, dtype=X.dtype), torch.eye(y, dtype=X.dtype)
    Cs = [c_scale * Bd[:, i:i + 1] @ eye_u[i:i + 1] for i in range(u)]
    Ds = [d_scale * eye_y[:, j:j + 1] @ Fd[j:j + 1] for j in range(y)]
    Dm = Fd @ Bd - Fa @ Ba
    n = x + b
    ll = torch.zeros(N, dtype=X.dtype)

    def joint(t):
        Lt, Kt = L[t], K[t]
        Fj = torch.cat([torch.cat([Ad, Bd @ Lt], -1), torch.cat([Kt @ Fd @ Ad, Aa + Ba @ Lt - Kt @ Fa @ Aa + Kt @ Dm @ Lt], -1)], -2)
        Gj = torch.cat([torch.cat([Vd, torch.zeros(x, y, dtype=X.dtype)], -1), torch.cat([Kt @ Fd @ Vd, Kt @ Wd], -1)], -2)
        return Fj, Gj @ mT(Gj)

    _, N0 = joint(0)
    for i in range(N):                                                   # the covariance is per trial here
        mu = torch.cat([X[i, 0], torch.zeros(n - d, dtype=X.dtype)])
        Sig = N0
        for t in range(T):
            Fj, Nj = joint(t)
            Soo = Sig[:d, :d]
            g = torch.linalg.solve(Soo, X[i, t] - mu[:d])
            mu_c = mu + Sig[:, :d] @ g
            Sig_c = Sig - Sig[:, :d] @ torch.linalg.solve(Soo, Sig[:d, :])
            Z2 = Sig_c + torch.outer(mu_c, mu_c)
            U2 = L[t] @ Z2[x:, x:] @ mT(L[t])
            mu = Fj @ mu_c
            Sig = Fj @ Sig_c @ mT(Fj) + Nj
            for Ci in Cs:
                gi = torch.cat([Ci, K[t] @ Fd @ Ci], 0)
                Sig = Sig + gi @ U2 @ mT(gi)
            X2 = Sig[:x, :x] + torch.outer(mu[:x], mu[:x])
            add = sum((K[t] @ Dj) @ X2 @ mT(K[t] @ Dj) for Dj in Ds)
            Sig = Sig + torch.nn.functional.pad(add, (x, 0, x, 0))
            Sig = 0.5 * (Sig + mT(Sig))
            Lc = torch.linalg.cholesky(Sig[:d, :d])
            z = torch.linalg.solve_triangular(Lc, (X[i, t + 1] - mu[:d])[:, None], upper=False)[:, 0]
            ll[i] = ll[i] - 0.5 * d * LOG2PI - torch.log(torch.diagonal(Lc)).sum() - 0.5 * (z * z).sum()
    return ll

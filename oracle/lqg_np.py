"""Float64 NumPy restatement of the reference likelihood path (TEST ORACLE ONLY).

Every function cites the reference file:line it follows (paths relative to
/root/reference).  Written from the mathematics, not copied; see
``oracle/__init__.py`` for the parity status ("unpinned by the reference's own
tests", pinned by invariants).

Conventions: a *spec* is a dict with time-stacked float64 arrays
``A[T,b,b] B[T,b,u] F[T,y,b] V[T,b,b] W[T,y,y] Q[T,b,b] R[T,u,u]`` plus the
fill-ins of ``lqg/utils.py:26-35`` (``q, Qf, qf, P, r``).
"""
from __future__ import annotations

import math
from itertools import chain

import numpy as np
import scipy.linalg as sla

LOG2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------- specs
def time_stack_spec(A, B, F, V, W, Q, R, T):
    """lqg/utils.py:10-35 -- stack T copies, zero affine terms, Qf = Q[-1]."""
    st = lambda M: np.broadcast_to(np.asarray(M, dtype=np.float64), (T,) + np.shape(M)).copy()
    A, B, F, V, W, Q, R = map(st, (A, B, F, V, W, Q, R))
    b, u = Q.shape[1], R.shape[1]
    return dict(A=A, B=B, F=F, V=V, W=W, Q=Q, R=R,
                q=np.zeros((T, b)), Qf=Q[-1].copy(), qf=np.zeros(b),
                P=np.zeros((T, u, b)), r=np.zeros((T, u)))


def dynamics_spec(A, B, F, V, W, T):
    """lqg/system.py:331-345 -- Dynamics(): zero cost matrices."""
    x, u = A.shape[0], B.shape[1]
    return time_stack_spec(A, B, F, V, W, np.zeros((x, x)), np.zeros((u, u)), T)


# --------------------------------------------------------------------------- gains
def lqr_backward(spec, eps=1e-8):
    """lqg/control/lqr.py:16-42 -- reverse Riccati sweep.  Returns L[T,u,b], l[T,u], H[T,u,u]."""
    T = spec["A"].shape[0]
    S, s = spec["Qf"].copy(), spec["qf"].copy()
    u, b = spec["B"].shape[2], spec["A"].shape[1]
    Ls, ls, Hs = np.zeros((T, u, b)), np.zeros((T, u)), np.zeros((T, u, u))
    for t in range(T - 1, -1, -1):
        Q, q, P, R, r, A, B = (spec[k][t] for k in ("Q", "q", "P", "R", "r", "A", "B"))
        H = R + B.T @ S @ B                                   # lqr.py:22
        G = P + B.T @ S @ A                                   # lqr.py:23
        g = r + B.T @ s                                       # lqr.py:24
        lam_min = np.linalg.eigvalsh(0.5 * (H + H.T))[0]      # lqr.py:27 (eigh symmetrises)
        Ht = H + max(0.0, eps - lam_min) * np.eye(u)          # lqr.py:28
        L = -np.linalg.solve(Ht, G)                           # lqr.py:30
        l = -np.linalg.solve(Ht, g)                           # lqr.py:31
        S_new = Q + A.T @ S @ A + L.T @ H @ L + L.T @ G + G.T @ L     # lqr.py:33 (un-shifted H)
        s = q + A.T @ s + G.T @ l + L.T @ H @ l + L.T @ g             # lqr.py:34
        S = S_new
        Ls[t], ls[t], Hs[t] = L, l, Ht
    return Ls, ls, Hs


def kf_forward(spec, Sigma0):
    """lqg/belief/kf.py:6-21 -- forward Kalman-gain sweep.  Returns K[T,b,y]."""
    T = spec["A"].shape[0]
    b, y = spec["A"].shape[1], spec["F"].shape[1]
    P = np.array(Sigma0, dtype=np.float64)
    Ks = np.zeros((T, b, y))
    for t in range(T):
        A, F, V, W = (spec[k][t] for k in ("A", "F", "V", "W"))
        P = A @ P @ A.T + V @ V.T                             # kf.py:10
        G = F @ P @ F.T + W @ W.T                             # kf.py:11
        K = P @ F.T @ np.linalg.inv(G)                        # kf.py:12
        P = (np.eye(b) - K @ F) @ P                           # kf.py:14
        Ks[t] = K
    return Ks


# --------------------------------------------------------------------------- joint system
def joint_system(actor, dyn, L, K):
    """lqg/system.py:163-207 -- joint (x, xhat) transition F[T,n,n] and noise factor G[T,n,x+y]."""
    Ad, Bd, Fd, Vd, Wd = (dyn[k] for k in ("A", "B", "F", "V", "W"))
    Aa, Ba, Fa = (actor[k] for k in ("A", "B", "F"))
    T, x = Ad.shape[0], Ad.shape[1]
    y = Wd.shape[2]
    top = np.concatenate([Ad, Bd @ L], axis=-1)
    bot = np.concatenate([K @ Fd @ Ad,
                          Aa + Ba @ L - K @ Fa @ Aa + K @ (Fd @ Bd - Fa @ Ba) @ L], axis=-1)
    Fj = np.concatenate([top, bot], axis=-2)
    Gj = np.concatenate([np.concatenate([Vd, np.zeros((T, x, y))], axis=-1),
                         np.concatenate([K @ Fd @ Vd, K @ Wd], axis=-1)], axis=-2)
    return Fj, Gj


def conditional_moments(actor, dyn, xs, Sigma0=None):
    """lqg/system.py:142-235 -- predictive moments of one trial ``xs[T+1,d]``.

    Returns mu[T,n], Sigma[T,n,n]; index t holds the moments of (x, xhat)_{t+1} given x_{0..t}.
    """
    xs = np.asarray(xs, dtype=np.float64)
    d = xs.shape[1]
    x, b = dyn["A"].shape[1], actor["A"].shape[1]
    L, _, _ = lqr_backward(actor)                                               # system.py:157
    K = kf_forward(actor, actor["V"][0] @ actor["V"][0].T if Sigma0 is None else Sigma0)  # :158-161
    Fj, Gj = joint_system(actor, dyn, L, K)
    T = Fj.shape[0]
    assert xs.shape[0] == T + 1, "need T+1 observations (SURVEY H7)"
    mu = np.concatenate([xs[0], np.zeros(x - d + b)])                            # system.py:211
    Sig = Gj[0] @ Gj[0].T                                                        # system.py:212
    mus, Sigs = np.zeros((T, x + b)), np.zeros((T, x + b, x + b))
    for t in range(T):
        F, G = Fj[t], Gj[t]
        FS = F @ Sig
        Soo = Sig[:d, :d]
        mu = F @ mu + FS[:, :d] @ np.linalg.solve(Soo, xs[t] - mu[:d])           # system.py:219-221
        Sig = F @ Sig @ F.T + G @ G.T - FS[:, :d] @ np.linalg.solve(Soo, (Sig @ F.T)[:d, :])  # :223-230
        mus[t], Sigs[t] = mu, Sig
    return mus, Sigs


def mvn_logpdf(xv, mu, Sigma):
    """numpyro MultivariateNormal.log_prob: Cholesky, triangular solve, log-diag sum."""
    Lc = np.linalg.cholesky(Sigma)
    z = sla.solve_triangular(Lc, xv - mu, lower=True)
    return -0.5 * len(xv) * LOG2PI - np.log(np.diag(Lc)).sum() - 0.5 * z @ z


def log_likelihood(actor, dyn, X, Sigma0=None):
    """lqg/system.py:237-248 -- per-trial log-likelihood of ``X[N,T+1,d]`` (sum over time)."""
    X = np.asarray(X, dtype=np.float64)
    N, _, d = X.shape
    out = np.zeros(N)
    for i in range(N):
        mus, Sigs = conditional_moments(actor, dyn, X[i], Sigma0)
        out[i] = sum(mvn_logpdf(X[i, t + 1], mus[t, :d], Sigs[t, :d, :d]) for t in range(mus.shape[0]))
    return out


# --------------------------------------------------------------------------- data generator
def simulate(actor, dyn, n, rng, x0=None, xhat0=None, Sigma0=None, return_all=False):
    """lqg/system.py:62-140 with a NumPy Generator instead of JAX threefry (so samples differ
    from the reference's, the distribution does not).  Returns x[n,T+1,xdim]."""
    T, xd, bd, yd = dyn["A"].shape[0], dyn["A"].shape[1], actor["A"].shape[1], dyn["F"].shape[1]
    L, l, _ = lqr_backward(actor)
    K = kf_forward(actor, actor["V"][0] @ actor["V"][0].T if Sigma0 is None else Sigma0)
    xs = np.zeros((n, T + 1, xd)); xh = np.zeros((n, T + 1, bd))
    ys = np.zeros((n, T, yd)); us = np.zeros((n, T, l.shape[1]))
    if x0 is not None:
        xs[:, 0] = x0
    if xhat0 is not None:
        xh[:, 0] = xhat0
    eps = rng.standard_normal((n, T, xd)); eta = rng.standard_normal((n, T, yd))
    for t in range(T):
        u = xh[:, t] @ L[t].T + l[t]                                              # system.py:110
        xn = xs[:, t] @ dyn["A"][t].T + u @ dyn["B"][t].T + eps[:, t] @ dyn["V"][t].T   # :113-117
        yv = xn @ dyn["F"][t].T + eta[:, t] @ dyn["W"][t].T                       # :120
        xp = xh[:, t] @ actor["A"][t].T + u @ actor["B"][t].T                     # :123
        xh[:, t + 1] = xp + (yv - xp @ actor["F"][t].T) @ K[t].T                  # :124
        xs[:, t + 1], ys[:, t], us[:, t] = xn, yv, u
    return (xs, xh, ys, us) if return_all else xs


# --------------------------------------------------------------------------- model zoo (base matrices)
def _bd(block, dim):
    return sla.block_diag(*[np.asarray(block, dtype=np.float64)] * dim)


def bounded_actor_mats(dim=1, process_noise=1.0, action_variability=0.5, sigma_target=6.0,
                       sigma_cursor=6.0, action_cost=1.0, dt=1.0 / 60.0):
    """lqg/tracking/basic.py:7-64 (TrackingTask / BoundedActor)."""
    d = 2 * dim
    A = np.eye(d); B = dt * _bd([[0.0], [1.0]], dim); F = np.eye(d)
    V = np.diag([process_noise, action_variability] * dim)
    W = np.diag([sigma_target, sigma_cursor] * dim)
    Q = _bd([[1.0, -1.0], [-1.0, 1.0]], dim); R = np.eye(dim) * action_cost
    act = dict(A=A, B=B, F=F, V=V, W=W, Q=Q, R=R)
    return act, dict(act)


def optimal_actor_mats(**kw):
    """lqg/tracking/basic.py:67-87."""
    return bounded_actor_mats(action_cost=1e-3, **kw)


def relative_observation_mats(dim=1, process_noise=1.0, action_variability=0.5, sigma=6.0,
                              action_cost=1.0, dt=1.0 / 60.0):
    """lqg/tracking/basic.py:90-124."""
    act, _ = bounded_actor_mats(dim=dim, process_noise=process_noise, action_variability=action_variability,
                                action_cost=action_cost, dt=dt)
    act["F"] = _bd([[1.0, -1.0]], dim); act["W"] = np.diag([sigma] * dim)
    return act, dict(act)


def swap_dims(d, dim):
    """lqg/tracking/subjective.py:7-12 -- observed (target, cursor) pairs first, the rest last."""
    idx = list(range(d)); k = d // dim
    obs = [idx[k * i:k * i + 2] for i in range(dim)]
    un = [idx[k * i + 2:k * (i + 1)] for i in range(dim)]
    return list(chain(*(obs + un)))


def subjective_actor_mats(dim=1, process_noise=1.0, action_cost=1.0, action_variability=0.5,
                          subj_noise=1.0, subj_vel_noise=0.5, sigma_target=6.0, sigma_cursor=6.0,
                          dt=1.0 / 60.0):
    """lqg/tracking/subjective.py:15-47."""
    dyn = dict(A=np.eye(2 * dim), B=_bd([[0.0], [dt]], dim), F=np.eye(2 * dim),
               V=_bd(np.diag([process_noise, action_variability]), dim),
               W=_bd(np.diag([sigma_target, sigma_cursor]), dim))
    dyn["Q"] = np.zeros((2 * dim, 2 * dim)); dyn["R"] = np.zeros((dim, dim))
    A = _bd([[1.0, 0.0, dt], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]], dim)
    B = _bd([[0.0], [dt], [0.0]], dim)
    F = _bd([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], dim)
    V = _bd(np.diag([subj_noise, action_variability, subj_vel_noise]), dim)
    Q = _bd([[1.0, -1.0, 0.0], [-1.0, 1.0, 0.0], [0.0, 0.0, 0.0]], dim)
    R = np.eye(dim) * action_cost
    p = swap_dims(3 * dim, dim)
    act = dict(A=A[p][:, p], B=B[p], F=F[:, p], V=V[p], W=dyn["W"].copy(), Q=Q[p][:, p], R=R)
    return act, dyn


def _make_psd(M, eps=1e-6):
    """lqg/tracking/point_mass.py:128-144."""
    w, U = np.linalg.eigh(0.5 * (M + M.T))
    return U @ np.diag(np.clip(w, eps, None)) @ U.T


def point_mass_mats(process_noise=1.0, action_variability=1e-3, sigma_target=6.0, sigma_cursor=6.0,
                    action_cost=0.01, dt=1.0 / 60.0, damping=0.1, m=1.0, tau=0.0015):
    """lqg/tracking/point_mass.py:7-125 (ZOH via expm, Van-Loan noise, *upper* Cholesky factor)."""
    Ac = np.array([[0.0, 1.0, 0.0], [0.0, -damping / m, 1.0 / m], [0.0, 0.0, -1.0 / tau]])
    Bc = np.array([[0.0], [0.0], [1.0 / tau]])
    M = np.zeros((4, 4)); M[:3, :3] = Ac; M[:3, 3:] = Bc
    E = sla.expm(M * dt); Ad, Bdm = E[:3, :3], E[:3, 3:]                       # :50-79
    Gn = 1e-2 * action_variability * Bc
    VL = np.block([[Ac, Gn @ Gn.T], [np.zeros((3, 3)), -Ac.T]])
    Qd = sla.expm(VL * dt)[:3, 3:]                                              # :82-110 (as written there)
    Vn = sla.cholesky(_make_psd(Qd), lower=False)                               # :123-125 upper factor
    A = sla.block_diag(np.eye(1), Ad); B = np.vstack([np.zeros((1, 1)), Bdm])
    V = sla.block_diag(np.diag([process_noise]), Vn)
    F = np.eye(3, 4); W = np.diag([sigma_target, sigma_cursor, sigma_cursor])
    Q = np.zeros((4, 4)); Q[:2, :2] = [[1.0, -1.0], [-1.0, 1.0]]
    R = np.eye(1) * action_cost * dt
    act = dict(A=A, B=B, F=F, V=V, W=W, Q=Q, R=R)
    return act, dict(act)


def hand_model_mats(process_noise=1.0, action_variability=0.5, sigma_target=6.0, sigma_cursor=6.0, action_cost=1.0,
                    dt=1.0 / 60.0, m=1.0, tau=0.04, position_noise=0.0):
    """notebooks/HandModel.ipynb cell 2 (HandMotionModelTrackingTask); position_noise is this repo's regulariser."""
    A = np.zeros((5, 5)); A[0, 0] = 1.0
    A[1:, 1:] = [[1.0, dt, 0.0, 0.0], [0.0, 1.0, dt / m, 0.0], [0.0, 0.0, 1.0 - dt / tau, dt / tau], [0.0, 0.0, 0.0, 1.0 - dt / tau]]
    B = np.zeros((5, 1)); B[4, 0] = dt / tau
    F = np.eye(2, 5)
    V = np.diag([process_noise, position_noise, 0.0, 0.0, action_variability])
    W = np.diag([sigma_target, sigma_cursor])
    Q = np.zeros((5, 5)); Q[:2, :2] = [[1.0, -1.0], [-1.0, 1.0]]
    R = np.eye(1) * action_cost
    act = dict(A=A, B=B, F=F, V=V, W=W, Q=Q, R=R)
    return act, dict(act)


def delay_mats(mats, delay):
    """lqg/tracking/delay.py:9-33 applied to base matrices (shift-register augmentation)."""
    A, B, F, V, W, Q, R = (mats[k] for k in ("A", "B", "F", "V", "W", "Q", "R"))
    d = A.shape[0]
    A2 = sla.block_diag(A, np.zeros((d * delay, d * delay))) + np.diag(np.ones(d * delay), k=-d)
    B2 = np.vstack([B] + [np.zeros_like(B)] * delay)
    F2 = np.hstack([np.zeros((F.shape[0], F.shape[1] * delay)), F])
    V2 = sla.block_diag(V, np.zeros((d * delay, d * delay)))
    Q2 = sla.block_diag(Q, *[np.zeros_like(Q)] * delay)
    return dict(A=A2, B=B2, F=F2, V=V2, W=W.copy(), Q=Q2, R=R.copy())


def make_system(mats_pair, T):
    """(actor_mats, dyn_mats) -> (actor_spec, dyn_spec), cf. System.__init__ (lqg/system.py:12-15)."""
    act, dyn = mats_pair
    return (time_stack_spec(*(act[k] for k in ("A", "B", "F", "V", "W", "Q", "R")), T),
            time_stack_spec(*(dyn[k] for k in ("A", "B", "F", "V", "W", "Q", "R")), T))

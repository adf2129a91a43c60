"""Signal-dependent-noise LQG (Todorov 2005, "Stochastic optimal control and estimation methods adapted to the noise
characteristics of the sensorimotor system", Neural Computation 17) in float64 NumPy -- TEST ORACLE ONLY.

NOT in the reference: RothkopfLab/lqg has no signal-dependent noise (docs/README.md:60-62 lists it as future work; only
lqg/infer/prior.py:11 mentions the name), so there is no reference output to compare with -- PARITY UNPINNED.  The north star
asks for "Todorov-style alternating control/estimator gain iterations under signal-dependent noise" as one kernel; this file
restates the published algorithm and `tests/test_sdn.py` pins it by Monte Carlo simulation of the
controlled system and by its reduction to plain LQR / Kalman gains when the multiplicative noise vanishes.

Model (Todorov's predictor-form convention; the control acts on the prediction x̂_t):
    x_{t+1} = A x_t + B u_t + xi_t + sum_i eps^i_t C_i u_t         xi ~ N(0, Om_xi),  eps^i ~ N(0, 1)
    y_t     = H x_t + om_t + sum_i eta^i_t D_i x_t                  om ~ N(0, Om_om),  eta^i ~ N(0, 1)
    x̂_{t+1} = A x̂_t + B u_t + K_t (y_t - H x̂_t)                   u_t = -L_t x̂_t
    cost    = sum_{t<T} (x_t' Q x_t + u_t' R u_t) + x_T' Q_f x_T
Given the filter gains K the optimal control gains L follow from a backward pass, given L the optimal K from a forward pass
over the second moments of e = x - x̂ and x̂; the two are alternated until they stop changing.
"""
from __future__ import annotations

import numpy as np


def backward_pass(A, B, H, C, D, Q, R, Qf, Om_xi, Om_om, K):
    """K[T,b,y] -> L[T,u,b], (Sx_0, Se_0, s_0).  Todorov (2005) eq. 4.2."""
    T = K.shape[0]
    b, u = B.shape
    L = np.zeros((T, u, b))
    Sx, Se, s = Qf.copy(), np.zeros((b, b)), 0.0
    for t in range(T - 1, -1, -1):
        Kt = K[t]
        M = R + B.T @ Sx @ B + sum(Ci.T @ (Sx + Se) @ Ci for Ci in C)
        L[t] = np.linalg.solve(M, B.T @ Sx @ A)
        AKH = A - Kt @ H
        s = np.trace(Sx @ Om_xi + Se @ (Om_xi + Kt @ Om_om @ Kt.T)) + s
        Sx_new = Q + A.T @ Sx @ (A - B @ L[t]) + sum(Di.T @ Kt.T @ Se @ Kt @ Di for Di in D)
        Se = A.T @ Sx @ B @ L[t] + AKH.T @ Se @ AKH
        Sx = 0.5 * (Sx_new + Sx_new.T)
        Se = 0.5 * (Se + Se.T)
    return L, (Sx, Se, s)


def forward_pass(A, B, H, C, D, Om_xi, Om_om, Sigma1, xhat1, L):
    """L[T,u,b] -> K[T,b,y].  Todorov (2005) eq. 5.2 (no internal estimator noise)."""
    T = L.shape[0]
    b, y = A.shape[0], H.shape[0]
    K = np.zeros((T, b, y))
    Se, Sx, Sxe = Sigma1.copy(), np.outer(xhat1, xhat1), np.zeros((b, b))
    for t in range(T):
        Lt = L[t]
        tot = Se + Sx + Sxe + Sxe.T
        G = H @ Se @ H.T + Om_om + sum(Di @ tot @ Di.T for Di in D)
        K[t] = A @ Se @ H.T @ np.linalg.inv(G)
        ABL, AKH = A - B @ Lt, A - K[t] @ H
        Se_new = Om_xi + AKH @ Se @ A.T + sum(Ci @ Lt @ Sx @ Lt.T @ Ci.T for Ci in C)
        Sx_new = K[t] @ H @ Se @ A.T + ABL @ Sx @ ABL.T + ABL @ Sxe @ H.T @ K[t].T + K[t] @ H @ Sxe.T @ ABL.T
        Sxe = ABL @ Sxe @ AKH.T
        Se, Sx = 0.5 * (Se_new + Se_new.T), 0.5 * (Sx_new + Sx_new.T)
    return K


def solve(A, B, H, C, D, Q, R, Qf, Om_xi, Om_om, Sigma1, xhat1, T, sweeps=10):
    """Alternating iterations starting from K = 0.  Returns L[T,u,b], K[T,b,y], expected cost."""
    b, y = A.shape[0], H.shape[0]
    K = np.zeros((T, b, y))
    for _ in range(sweeps):
        L, (Sx, Se, s) = backward_pass(A, B, H, C, D, Q, R, Qf, Om_xi, Om_om, K)
        K = forward_pass(A, B, H, C, D, Om_xi, Om_om, Sigma1, xhat1, L)
    L, (Sx, Se, s) = backward_pass(A, B, H, C, D, Q, R, Qf, Om_xi, Om_om, K)
    cost = float(xhat1 @ Sx @ xhat1 + np.trace((Sx + Se) @ Sigma1) + s)
    return L, K, cost


def simulate_cost(A, B, H, C, D, Q, R, Qf, Om_xi, Om_om, Sigma1, xhat1, L, K, n, rng):
    """Monte Carlo estimate (mean, standard error) of the total cost of the closed loop with gains L, K."""
    T = L.shape[0]
    b, y = A.shape[0], H.shape[0]
    cxi, com, c1 = (np.linalg.cholesky(M + 1e-300 * np.eye(M.shape[0])) if np.any(M) else np.zeros_like(M) for M in (Om_xi, Om_om, Sigma1))
    xh = np.tile(xhat1, (n, 1))
    x = xh + rng.standard_normal((n, b)) @ c1.T
    cost = np.zeros(n)
    for t in range(T):
        u = -xh @ L[t].T
        cost += np.einsum("ni,ij,nj->n", x, Q, x) + np.einsum("ni,ij,nj->n", u, R, u)
        yv = x @ H.T + rng.standard_normal((n, y)) @ com.T
        for Di in D:
            yv = yv + rng.standard_normal((n, 1)) * (x @ Di.T)
        xn = x @ A.T + u @ B.T + rng.standard_normal((n, b)) @ cxi.T
        for Ci in C:
            xn = xn + rng.standard_normal((n, 1)) * (u @ Ci.T)
        xh = xh @ A.T + u @ B.T + (yv - xh @ H.T) @ K[t].T
        x = xn
    cost += np.einsum("ni,ij,nj->n", x, Qf, x)
    return float(cost.mean()), float(cost.std() / np.sqrt(n))


def tracking_example(dt=1.0 / 60.0, sigma_u=0.5, sigma_target=6.0, sigma_cursor=6.0, c_mult=0.5, d_mult=0.2, action_cost=1.0):
    """A signal-dependent-noise variant of the reference's 1-D tracking task (lqg/tracking/basic.py:7-41): target random walk
    + cursor integrator, control-dependent motor noise (C u) and state-dependent observation noise on the cursor (D x)."""
    A = np.eye(2)
    B = np.array([[0.0], [dt]])
    H = np.eye(2)
    C = [np.array([[0.0], [c_mult * dt]])]
    D = [np.array([[0.0, 0.0], [0.0, d_mult]])]
    Q = np.array([[1.0, -1.0], [-1.0, 1.0]])
    R = np.array([[action_cost]])
    Om_xi = np.diag([1.0, sigma_u ** 2]) * 1.0
    Om_om = np.diag([sigma_target ** 2, sigma_cursor ** 2])
    Sigma1 = np.diag([1.0, 1.0])
    xhat1 = np.array([0.5, -0.5])
    return dict(A=A, B=B, H=H, C=C, D=D, Q=Q, R=R, Qf=Q, Om_xi=Om_xi, Om_om=Om_om, Sigma1=Sigma1, xhat1=xhat1)


# =========================================================================================== likelihood under signal-dependent noise
# Extension of the reference's experimenter-side filter (lqg/system.py:142-248) -- NOT in the reference, PARITY UNPINNED.
# Written in the reference's own FILTER-form conventions (lqg/system.py:110-124, lqg/belief/kf.py:10-14) so that with C = D = []
# every function below reproduces oracle/lqg_np.py (= the reference) exactly:
#     u_t      = L_t xhat_t (+ l_t)
#     x_{t+1}  = A x_t + B u_t + V eps_t          + sum_i eps'_{i,t} C_i u_t          C_i [x, u]   control-dependent noise
#     y_t      = F x_{t+1} + W eta_t              + sum_j eta'_{j,t} D_j x_{t+1}      D_j [y, x]   state-dependent observation noise
#     xp       = A_a xhat_t + B_a u_t ;  xhat_{t+1} = xp + K_t (y_t - F_a xp)
# The gains (L, K) are INPUTS (any source: the actor's lqr.backward / kf.forward as in the reference, or an SDN-aware solver).
# With multiplicative noise the joint (x, xhat) process is no longer Gaussian; its first two moments still obey a closed
# recursion (the noise covariance depends on the second moment of the state), which `sdn_moments` propagates exactly.  The
# likelihood conditions on the observed part of x step by step like system.py:214-235 and treats each predictive distribution as
# Gaussian with those moments (moment matching) -- the approximation of Schultheis et al. (NeurIPS 2021) the reference's README
# (docs/README.md:60-62) points to.  The covariance now depends on the running mean, so it is per TRIAL, not per parameter set.
from oracle import lqg_np as _O  # noqa: E402


def sdn_joint_terms(dyn, K, C, D, t):
    """g_i = [C_i ; K_t F_d C_i]  (n x u)  and  h_j = K_t D_j  (b x x)."""
    Kt, Fd = K[t], dyn["F"][t]
    return [np.vstack([Ci, Kt @ Fd @ Ci]) for Ci in C], [Kt @ Dj for Dj in D]


def sdn_predict(F, N, L, gs, hs, mu_c, Sig_c, x):
    """One prediction step from the (conditioned or not) moments (mu_c, Sig_c) of z_t = (x, xhat)_t."""
    Z2 = Sig_c + np.outer(mu_c, mu_c)                              # E[z z^T]
    U2 = L @ Z2[x:, x:] @ L.T                                      # E[u u^T]
    mu = F @ mu_c
    Sig = F @ Sig_c @ F.T + N + sum(g @ U2 @ g.T for g in gs)
    X2 = Sig[:x, :x] + np.outer(mu[:x], mu[:x])                    # E[x' x'^T] given the same information
    if hs:
        Sig = Sig.copy()
        Sig[x:, x:] += sum(h @ X2 @ h.T for h in hs)
    return mu, 0.5 * (Sig + Sig.T)


def sdn_moments(actor, dyn, L, K, C, D, x0=None, xhat0=None):
    """EXACT unconditional mean[T+1,n] / covariance[T+1,n,n] of z_t = (x_t, xhat_t) under the generative model above."""
    Fj, Gj = _O.joint_system(actor, dyn, L, K)
    T, x, b = Fj.shape[0], dyn["A"].shape[1], actor["A"].shape[1]
    mu = np.concatenate([np.zeros(x) if x0 is None else x0, np.zeros(b) if xhat0 is None else xhat0])
    Sig = np.zeros((x + b, x + b))
    mus, Sigs = [mu], [Sig]
    for t in range(T):
        gs, hs = sdn_joint_terms(dyn, K, C, D, t)
        mu, Sig = sdn_predict(Fj[t], Gj[t] @ Gj[t].T, L[t], gs, hs, mu, Sig, x)
        mus.append(mu); Sigs.append(Sig)
    return np.stack(mus), np.stack(Sigs)


def sdn_conditional_moments(actor, dyn, L, K, C, D, xs):
    """system.py:142-235 with the signal-dependent terms: predictive moments mu[T,n], Sigma[T,n,n] of one trial xs[T+1,d]."""
    xs = np.asarray(xs, dtype=np.float64)
    d, x, b = xs.shape[1], dyn["A"].shape[1], actor["A"].shape[1]
    Fj, Gj = _O.joint_system(actor, dyn, L, K)
    T = Fj.shape[0]
    mu = np.concatenate([xs[0], np.zeros(x - d + b)])                             # system.py:211
    Sig = Gj[0] @ Gj[0].T                                                          # system.py:212
    mus, Sigs = np.zeros((T, x + b)), np.zeros((T, x + b, x + b))
    for t in range(T):
        Soo = Sig[:d, :d]
        mu_c = mu + Sig[:, :d] @ np.linalg.solve(Soo, xs[t] - mu[:d])              # condition on x_t ...
        Sig_c = Sig - Sig[:, :d] @ np.linalg.solve(Soo, Sig[:d, :])
        gs, hs = sdn_joint_terms(dyn, K, C, D, t)
        mu, Sig = sdn_predict(Fj[t], Gj[t] @ Gj[t].T, L[t], gs, hs, mu_c, Sig_c, x)   # ... then predict (system.py:219-230)
        mus[t], Sigs[t] = mu, Sig
    return mus, Sigs


def sdn_log_likelihood(actor, dyn, L, K, C, D, X):
    """Per-trial log-likelihood of X[N,T+1,d] (sum over time), system.py:237-248 with moment-matched Gaussians."""
    X = np.asarray(X, dtype=np.float64)
    N, _, d = X.shape
    out = np.zeros(N)
    for i in range(N):
        mus, Sigs = sdn_conditional_moments(actor, dyn, L, K, C, D, X[i])
        out[i] = sum(_O.mvn_logpdf(X[i, t + 1], mus[t, :d], Sigs[t, :d, :d]) for t in range(mus.shape[0]))
    return out


def sdn_simulate(actor, dyn, L, K, C, D, n, rng, x0=None, xhat0=None, l=None):
    """The generative model above (system.py:62-140 plus the multiplicative terms).  Returns x[n,T+1,x], xhat[n,T+1,b]."""
    T, xd, bd, yd = dyn["A"].shape[0], dyn["A"].shape[1], actor["A"].shape[1], dyn["F"].shape[1]
    xs = np.zeros((n, T + 1, xd)); xh = np.zeros((n, T + 1, bd))
    if x0 is not None:
        xs[:, 0] = x0
    if xhat0 is not None:
        xh[:, 0] = xhat0
    for t in range(T):
        u = xh[:, t] @ L[t].T + (0.0 if l is None else l[t])
        xn = xs[:, t] @ dyn["A"][t].T + u @ dyn["B"][t].T + rng.standard_normal((n, xd)) @ dyn["V"][t].T
        for Ci in C:
            xn = xn + rng.standard_normal((n, 1)) * (u @ Ci.T)
        yv = xn @ dyn["F"][t].T + rng.standard_normal((n, yd)) @ dyn["W"][t].T
        for Dj in D:
            yv = yv + rng.standard_normal((n, 1)) * (xn @ Dj.T)
        xp = xh[:, t] @ actor["A"][t].T + u @ actor["B"][t].T
        xh[:, t + 1] = xp + (yv - xp @ actor["F"][t].T) @ K[t].T
        xs[:, t + 1] = xn
    return xs, xh


# =========================================================================================== gains under signal-dependent noise, FILTER form
# The alternating iterations of Todorov (2005) re-derived for the reference's own conventions (lqg/system.py:110-124,
# lqg/control/lqr.py:16-42, lqg/belief/kf.py:6-21), so that without multiplicative noise they ARE lqr.backward / kf.forward:
#     u_t = L_t xhat_t ;  x_{t+1} = A x_t + B u_t + xi + sum_i eps_i C_i u_t ;  y_{t+1} = F x_{t+1} + om + sum_j eta_j D_j x_{t+1}
#     xp = A xhat_t + B u_t ;  xhat_{t+1} = xp + K_t (y_{t+1} - F xp) ;  cost = sum_{t<T} (x'Qx + u'Ru) + x_T' Qf x_T
# (one model: the actor plans and filters with its own A, B, F, V, W, as lqr.backward(actor) / kf.forward(actor) do).
# With e = x - xhat, a = A e + xi + sum_i eps_i C_i u (prior error of x_{t+1}) and b = om + sum_j eta_j D_j x_{t+1}:
#     e_{t+1} = (I - K F) a - K b ,   cost-to-go  v_t(x, e) = x' Sx_t x + e' Se_t e + s_t   (exact for the given later gains):
#     St = Sx + sum_j D_j' K' Se K D_j ;  H = R + B' St B + sum_i C_i' (St + (I-KF)' Se (I-KF)) C_i ;  L = -H^-1 B' St A
#     Sx <- Q + A' St (A + B L) ;  Se <- A' St B H^-1 B' St A + ((I-KF)A)' Se (I-KF)A ;  s <- s + tr(St Om_xi) + tr(Se ((I-KF) Om_xi (I-KF)' + K Om_om K'))
#     forward:  P- = A Sig_e A' + Om_xi + sum_i C_i L Sig_xh L' C_i' ;  Reff = Om_om + sum_j D_j E[x'x''] D_j' ;  K = P- F' (F P- F' + Reff)^-1
def filter_backward_pass(A, B, F, C, D, Q, R, Qf, Om_xi, Om_om, K):
    """K[T,b,y] -> L[T,u,b] (reference sign: u = L xhat) and (Sx_0, Se_0, s_0)."""
    T = K.shape[0]
    b, u = B.shape
    L = np.zeros((T, u, b))
    Sx, Se, s = Qf.copy(), np.zeros((b, b)), 0.0
    I = np.eye(b)
    for t in range(T - 1, -1, -1):
        Kt = K[t]
        IKF = I - Kt @ F
        St = Sx + sum(Dj.T @ Kt.T @ Se @ Kt @ Dj for Dj in D)
        H = R + B.T @ St @ B + sum(Ci.T @ (St + IKF.T @ Se @ IKF) @ Ci for Ci in C)
        L[t] = -np.linalg.solve(H, B.T @ St @ A)
        s = s + np.trace(St @ Om_xi) + np.trace(Se @ (IKF @ Om_xi @ IKF.T + Kt @ Om_om @ Kt.T))
        Abar = IKF @ A
        Se_new = -A.T @ St @ B @ L[t] + Abar.T @ Se @ Abar
        Sx_new = Q + A.T @ St @ (A + B @ L[t])
        Sx, Se = 0.5 * (Sx_new + Sx_new.T), 0.5 * (Se_new + Se_new.T)
    return L, (Sx, Se, s)


def _filter_step_moments(A, B, F, C, D, Om_xi, Om_om, Lt, Se, Sx, Sxe):
    """Prior error covariance P-, effective observation-noise covariance Reff and E[x_{t+1} x_{t+1}'] from the moments of
    (xhat_t, e_t): Se = E[e e'], Sx = E[xhat xhat'], Sxe = E[xhat e']."""
    ABL = A + B @ Lt
    ctrl = sum(Ci @ Lt @ Sx @ Lt.T @ Ci.T for Ci in C) if C else 0.0
    Pm = A @ Se @ A.T + Om_xi + ctrl
    X2 = ABL @ Sx @ ABL.T + A @ Se @ A.T + ABL @ Sxe @ A.T + A @ Sxe.T @ ABL.T + Om_xi + ctrl
    Reff = Om_om + (sum(Dj @ X2 @ Dj.T for Dj in D) if D else 0.0)
    return ABL, Pm, Reff, X2


def _filter_advance_moments(A, F, ABL, Pm, Reff, Kt, Sx, Sxe):
    IKF = np.eye(A.shape[0]) - Kt @ F
    xa = Sxe @ A.T                                                   # E[xhat a']
    Se_n = IKF @ Pm @ IKF.T + Kt @ Reff @ Kt.T
    Sx_n = ABL @ Sx @ ABL.T + Kt @ F @ Pm @ F.T @ Kt.T + Kt @ Reff @ Kt.T + ABL @ xa @ F.T @ Kt.T + Kt @ F @ xa.T @ ABL.T
    Sxe_n = ABL @ xa @ IKF.T + Kt @ F @ Pm @ IKF.T - Kt @ Reff @ Kt.T
    return 0.5 * (Se_n + Se_n.T), 0.5 * (Sx_n + Sx_n.T), Sxe_n


def filter_forward_pass(A, B, F, C, D, Om_xi, Om_om, Sigma0, xhat0, L):
    """L[T,u,b] -> K[T,b,y] (filter form, kf.py:10-14 with the signal-dependent terms)."""
    T = L.shape[0]
    b, y = A.shape[0], F.shape[0]
    K = np.zeros((T, b, y))
    Se, Sx, Sxe = Sigma0.copy(), np.outer(xhat0, xhat0), np.zeros((b, b))
    for t in range(T):
        ABL, Pm, Reff, _ = _filter_step_moments(A, B, F, C, D, Om_xi, Om_om, L[t], Se, Sx, Sxe)
        K[t] = Pm @ F.T @ np.linalg.inv(F @ Pm @ F.T + Reff)
        Se, Sx, Sxe = _filter_advance_moments(A, F, ABL, Pm, Reff, K[t], Sx, Sxe)
    return K


def filter_solve(A, B, F, C, D, Q, R, Qf, Om_xi, Om_om, Sigma0, xhat0, T, sweeps=10):
    """Alternating iterations starting from K = 0.  Returns L[T,u,b], K[T,b,y], expected cost."""
    b, y = A.shape[0], F.shape[0]
    K = np.zeros((T, b, y))
    for _ in range(sweeps):
        L, _ = filter_backward_pass(A, B, F, C, D, Q, R, Qf, Om_xi, Om_om, K)
        K = filter_forward_pass(A, B, F, C, D, Om_xi, Om_om, Sigma0, xhat0, L)
    L, (Sx, Se, s) = filter_backward_pass(A, B, F, C, D, Q, R, Qf, Om_xi, Om_om, K)
    cost = float(xhat0 @ Sx @ xhat0 + np.trace((Sx + Se) @ Sigma0) + s)
    return L, K, cost


def filter_expected_cost(A, B, F, C, D, Q, R, Qf, Om_xi, Om_om, Sigma0, xhat0, L, K):
    """EXACT expected total cost of ARBITRARY gains (L, K) by propagating the second moments of (xhat, e)."""
    T = L.shape[0]
    Se, Sx, Sxe = Sigma0.copy(), np.outer(xhat0, xhat0), np.zeros_like(Sigma0)
    cost = 0.0
    for t in range(T):
        Exx = Sx + Se + Sxe + Sxe.T                                   # E[x x'], x = xhat + e
        cost += np.trace(Q @ Exx) + np.trace(L[t].T @ R @ L[t] @ Sx)
        ABL, Pm, Reff, _ = _filter_step_moments(A, B, F, C, D, Om_xi, Om_om, L[t], Se, Sx, Sxe)
        Se, Sx, Sxe = _filter_advance_moments(A, F, ABL, Pm, Reff, K[t], Sx, Sxe)
    return float(cost + np.trace(Qf @ (Sx + Se + Sxe + Sxe.T)))


def filter_simulate_cost(A, B, F, C, D, Q, R, Qf, Om_xi, Om_om, Sigma0, xhat0, L, K, n, rng):
    """Monte Carlo estimate (mean, standard error) of the total cost of the filter-form closed loop with gains L, K."""
    T = L.shape[0]
    b, y = A.shape[0], F.shape[0]
    chol = lambda M: np.linalg.cholesky(M + 1e-300 * np.eye(M.shape[0])) if np.any(M) else np.zeros_like(M)
    cxi, com, c0 = chol(Om_xi), chol(Om_om), chol(Sigma0)
    xh = np.tile(xhat0, (n, 1))
    x = xh + rng.standard_normal((n, b)) @ c0.T
    cost = np.zeros(n)
    for t in range(T):
        u = xh @ L[t].T
        cost += np.einsum("ni,ij,nj->n", x, Q, x) + np.einsum("ni,ij,nj->n", u, R, u)
        xn = x @ A.T + u @ B.T + rng.standard_normal((n, b)) @ cxi.T
        for Ci in C:
            xn = xn + rng.standard_normal((n, 1)) * (u @ Ci.T)
        yv = xn @ F.T + rng.standard_normal((n, y)) @ com.T
        for Dj in D:
            yv = yv + rng.standard_normal((n, 1)) * (xn @ Dj.T)
        xp = xh @ A.T + u @ B.T
        xh = xp + (yv - xp @ F.T) @ K[t].T
        x = xn
    cost += np.einsum("ni,ij,nj->n", x, Qf, x)
    return float(cost.mean()), float(cost.std() / np.sqrt(n))

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

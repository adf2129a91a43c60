"""Hand-derived reduced forward + reverse-mode equations in NumPy float64 (TEST ORACLE ONLY).

This is the *mathematical specification of the CUDA kernels* (lqg_b200/csrc): the same five forward
stages and their adjoints, for one parameter sample with time-invariant base matrices and N trials.
It is independent of the autodiff oracle (``oracle/lqg_torch.py``) and is checked against it in
``tests/test_oracle_invariants.py`` -- an "oracle of the oracle" (SURVEY 8c-v).

Reduced ("condition-then-predict") form, equivalent to reference lqg/system.py:214-235 (checked to
1e-15): with o = first d joint indices (observed), u = the other r = n-d,

    per sample : Sig' = F[:,u] C F[:,u]^T + G G^T ;  S' = Sig'[o,o] ;  J' = Sig'[u,o] S'^-1 ;
                 C <- Sig'[u,u] - J' S' J'^T                      (C_0 from Sig_0 = G_0 G_0^T)
    per trial  : e = x_{t+1} - F[o,o] x_t - F[o,u] c ;  c <- F[u,o] x_t + F[u,u] c + J' e   (c_0 = 0)
                 ll += -1/2 e^T S'^-1 e - 1/2 log|S'| - d/2 log 2pi

Derived per-sample constants (what the kernels keep in shared memory):
    FAd = Fd Ad, FAa = Fa Aa, D = Fd Bd - Fa Ba, N11 = Vd Vd^T, FN = Fd N11, Om = FN Fd^T + Wd Wd^T,
    VVa = Va Va^T, WWa = Wa Wa^T.
"""
from __future__ import annotations

import math

import numpy as np

LOG2PI = math.log(2.0 * math.pi)
sym = lambda M: 0.5 * (M + M.T)


def derive_consts(act, dyn):
    c = dict(Aa=act["A"], Ba=act["B"], Fa=act["F"], Q=sym(act["Q"]), R=sym(act["R"]),
             Ad=dyn["A"], Bd=dyn["B"], Fd=dyn["F"])
    c["VVa"] = act["V"] @ act["V"].T
    c["WWa"] = act["W"] @ act["W"].T
    c["FAd"] = dyn["F"] @ dyn["A"]
    c["FAa"] = act["F"] @ act["A"]
    c["D"] = dyn["F"] @ dyn["B"] - act["F"] @ act["B"]
    c["N11"] = dyn["V"] @ dyn["V"].T
    c["FN"] = dyn["F"] @ c["N11"]
    c["Om"] = c["FN"] @ dyn["F"].T + dyn["W"] @ dyn["W"].T
    return c


# ----------------------------------------------------------------------------------------- forward
def lqr_fwd(c, T, eps=1e-8):
    """Riccati sweep t = T-1..0 (lqr.py:16-42, q=r=P=0, symmetric S).  Returns L[T], Sric[T] (= S_{t+1})."""
    A, B, Q, R = c["Aa"], c["Ba"], c["Q"], c["R"]
    u, b = B.shape[1], A.shape[0]
    L = np.zeros((T, u, b)); Sric = np.zeros((T, b, b)); shift = np.zeros(T)
    S = Q.copy()
    for t in range(T - 1, -1, -1):
        Sric[t] = S
        SA, SB = S @ A, S @ B
        H = R + B.T @ SB
        G = B.T @ SA
        shift[t] = max(0.0, eps - np.linalg.eigvalsh(H)[0])
        Ht = H + shift[t] * np.eye(u)
        L[t] = -np.linalg.solve(Ht, G)
        S = sym(Q + A.T @ SA + L[t].T @ H @ L[t] + L[t].T @ G + G.T @ L[t])
    return L, Sric, shift


def kf_fwd(c, T, Sigma0=None):
    """Kalman-gain sweep (kf.py:6-21).  Returns K[T], Pkf[T] (= P before step t)."""
    A, F, VV, WW = c["Aa"], c["Fa"], c["VVa"], c["WWa"]
    b, y = A.shape[0], F.shape[0]
    K = np.zeros((T, b, y)); Pkf = np.zeros((T, b, b))
    P = VV.copy() if Sigma0 is None else np.array(Sigma0, dtype=np.float64)
    for t in range(T):
        Pkf[t] = P
        Pp = A @ P @ A.T + VV
        M = F @ Pp
        Gm = M @ F.T + WW
        K[t] = np.linalg.solve(Gm, M).T
        P = sym(Pp - K[t] @ M)
    return K, Pkf


def joint_F(c, L, K):
    top = np.hstack([c["Ad"], c["Bd"] @ L])
    bot = np.hstack([K @ c["FAd"], c["Aa"] + c["Ba"] @ L - K @ c["FAa"] + (K @ c["D"]) @ L])
    return np.vstack([top, bot])


def joint_N(c, K):
    KFN = K @ c["FN"]
    return np.block([[c["N11"], KFN.T], [KFN, K @ c["Om"] @ K.T]])


def _condition(Sig, d):
    """Sig (n x n) -> (chol(S), J = Sig[u,o] S^-1, C = Sig[u,u] - J S J^T)."""
    Lc = np.linalg.cholesky(Sig[:d, :d])
    Linv = np.linalg.inv(Lc)
    Z = Sig[d:, :d] @ Linv.T
    return Lc, Linv, Z @ Linv, sym(Sig[d:, d:] - Z @ Z.T)


def cov_fwd(c, L, K, d):
    """Per-sample covariance pass.  Returns C[T] (C_t before step t) and the per-step trial records."""
    T = L.shape[0]
    n = c["Ad"].shape[0] + c["Aa"].shape[0]
    r = n - d
    Cs = np.zeros((T, r, r))
    rec = dict(F=np.zeros((T, n, n)), J=np.zeros((T, r, d)), Linv=np.zeros((T, d, d)), logdet=np.zeros(T))
    _, _, J0, C = _condition(joint_N(c, K[0]), d)
    for t in range(T):
        Cs[t] = C
        Fj = joint_F(c, L[t], K[t])
        Fu = Fj[:, d:]
        Sig = Fu @ C @ Fu.T + joint_N(c, K[t])
        Lc, Linv, J, C = _condition(Sig, d)
        rec["F"][t], rec["J"][t], rec["Linv"][t] = Fj, J, Linv
        rec["logdet"][t] = np.log(np.diag(Lc)).sum()
    return Cs, rec, J0


def trial_fwd(rec, X):
    """Per-trial mean/likelihood pass.  X[N,T+1,d].  Returns ll[N], hist[T,N,r] (c_t before step t)."""
    N, T1, d = X.shape
    T = T1 - 1
    r = rec["J"].shape[1]
    c = np.zeros((N, r)); ll = np.zeros(N); hist = np.zeros((T, N, r))
    for t in range(T):
        hist[t] = c
        Fj, J, Linv = rec["F"][t], rec["J"][t], rec["Linv"][t]
        e = X[:, t + 1] - X[:, t] @ Fj[:d, :d].T - c @ Fj[:d, d:].T
        p = X[:, t] @ Fj[d:, :d].T + c @ Fj[d:, d:].T
        z = e @ Linv.T
        ll += -0.5 * (z * z).sum(1) - rec["logdet"][t] - 0.5 * d * LOG2PI
        c = p + e @ J.T
    return ll, hist


def forward(act, dyn, X, T=None, Sigma0=None):
    X = np.asarray(X, dtype=np.float64)
    T = X.shape[1] - 1 if T is None else T
    c = derive_consts(act, dyn)
    L, Sric, shift = lqr_fwd(c, T)
    K, Pkf = kf_fwd(c, T, Sigma0)
    Cs, rec, J0 = cov_fwd(c, L, K, X.shape[2])
    ll, hist = trial_fwd(rec, X)
    return ll, dict(c=c, L=L, Sric=Sric, K=K, Pkf=Pkf, Cs=Cs, rec=rec, hist=hist, J0=J0, shift=shift)


# ----------------------------------------------------------------------------------------- reverse
def trial_rev(rec, X, hist, w):
    """Per-trial adjoint, t = T-1..0, and the per-step sums over trials the sample adjoint consumes."""
    N, T1, d = X.shape
    T = T1 - 1
    n = rec["F"].shape[1]; r = n - d
    cb = np.zeros((N, r))
    sums = dict(Jb=np.zeros((T, r, d)), Fb=np.zeros((T, n, n)), Wv=np.zeros((T, d, d)))
    for t in range(T - 1, -1, -1):
        Fj, J, Linv = rec["F"][t], rec["J"][t], rec["Linv"][t]
        c = hist[t]
        e = X[:, t + 1] - X[:, t] @ Fj[:d, :d].T - c @ Fj[:d, d:].T
        v = (e @ Linv.T) @ Linv                              # S'^-1 e
        eb = cb @ J - w[:, None] * v
        sums["Jb"][t] = cb.T @ e
        sums["Fb"][t][d:, :d] = cb.T @ X[:, t]
        sums["Fb"][t][d:, d:] = cb.T @ c
        sums["Fb"][t][:d, :d] = -eb.T @ X[:, t]
        sums["Fb"][t][:d, d:] = -eb.T @ c
        sums["Wv"][t] = (w[:, None] * v).T @ v
        cb = cb @ Fj[d:, d:] - eb @ Fj[:d, d:]
    return sums


def _joint_bar(c, acc, L, K, Fb, Nb, Lb, Kb):
    """Push joint-level cotangents Fb (n x n) and symmetric Nb (n x n) of step t into the derived-constant
    accumulators and into Lb, Kb (restated on derived constants)."""
    x = c["Ad"].shape[0]
    F11, F12, F21, F22 = Fb[:x, :x], Fb[:x, x:], Fb[x:, :x], Fb[x:, x:]
    acc["Ad"] += F11
    acc["Bd"] += F12 @ L.T
    acc["FAd"] += K.T @ F21
    acc["Aa"] += F22
    acc["Ba"] += F22 @ L.T
    acc["FAa"] -= K.T @ F22
    acc["D"] += K.T @ F22 @ L.T
    Lb += c["Bd"].T @ F12 + (c["Ba"] + K @ c["D"]).T @ F22
    Kb += F21 @ c["FAd"].T - F22 @ c["FAa"].T + F22 @ (c["D"] @ L).T
    Nxx, Nbx, Nbb = Nb[:x, :x], Nb[x:, :x], Nb[x:, x:]
    acc["N11"] += Nxx
    acc["FN"] += 2.0 * K.T @ Nbx
    acc["Om"] += K.T @ Nbb @ K
    Kb += 2.0 * Nbx @ c["FN"].T + 2.0 * Nbb @ K @ c["Om"]


def cov_rev(c, L, K, Cs, sums, sw, d, acc):
    """Per-sample covariance adjoint, t = T-1..0.  Returns Lbar[T], Kbar[T]; fills ``acc``."""
    T = L.shape[0]
    n = c["Ad"].shape[0] + c["Aa"].shape[0]
    r = n - d
    Lbar = np.zeros_like(L); Kbar = np.zeros_like(K)
    Cb = np.zeros((r, r))
    for t in range(T - 1, -1, -1):
        Fj = joint_F(c, L[t], K[t]); Fu = Fj[:, d:]
        Sig = Fu @ Cs[t] @ Fu.T + joint_N(c, K[t])
        _, Linv, J, _ = _condition(Sig, d)
        Sinv = Linv.T @ Linv
        Jb = sums["Jb"][t]
        Sb = 0.5 * sums["Wv"][t] - 0.5 * sw * Sinv + J.T @ Cb @ J - J.T @ Jb @ Sinv
        Bb = -2.0 * Cb @ J + Jb @ Sinv
        Sgb = np.zeros((n, n))
        Sgb[:d, :d] = sym(Sb); Sgb[d:, :d] = 0.5 * Bb; Sgb[:d, d:] = 0.5 * Bb.T; Sgb[d:, d:] = Cb
        Fb = sums["Fb"][t].copy()
        Fb[:, d:] += 2.0 * Sgb @ Fu @ Cs[t]
        _joint_bar(c, acc, L[t], K[t], Fb, Sgb, Lbar[t], Kbar[t])
        Cb = sym(Fu.T @ Sgb @ Fu)
    # initial condition: C_0 = cond(N_0)
    _, _, J0, _ = _condition(joint_N(c, K[0]), d)
    Sgb = np.zeros((n, n))
    Sgb[d:, d:] = Cb; Sgb[d:, :d] = -Cb @ J0; Sgb[:d, d:] = Sgb[d:, :d].T; Sgb[:d, :d] = J0.T @ Cb @ J0
    _joint_bar(c, acc, L[0], K[0], np.zeros((n, n)), Sgb, Lbar[0], Kbar[0])
    return Lbar, Kbar


def kf_rev(c, K, Pkf, Kbar, acc, own_sigma0=True):
    """Kalman-gain adjoint, t = T-1..0 (SURVEY Appendix A.3 with symmetric cotangents)."""
    A, F, VV, WW = c["Aa"], c["Fa"], c["VVa"], c["WWa"]
    T, b = K.shape[0], A.shape[0]
    Pnb = np.zeros((b, b))
    for t in range(T - 1, -1, -1):
        P = Pkf[t]
        Pp = A @ P @ A.T + VV
        M = F @ Pp
        Gm = M @ F.T + WW
        Gi = np.linalg.inv(Gm)
        Kt = M.T @ Gi
        Ppb = Pnb.copy()
        Ktot = Kbar[t] - Pnb @ M.T
        Mb = -Kt.T @ Pnb
        Y = Ktot @ Gi
        Mb = Mb + Y.T
        Gmb = -sym(Kt.T @ Y)
        acc["Fa"] += Mb @ Pp + 2.0 * Gmb @ M
        Ppb += sym(F.T @ Mb) + F.T @ Gmb @ F
        acc["WWa"] += Gmb
        acc["Aa"] += 2.0 * Ppb @ A @ P
        acc["VVa"] += Ppb
        Pnb = A.T @ Ppb @ A
    if own_sigma0:
        acc["VVa"] += Pnb
    return Pnb


def lqr_rev(c, L, Sric, Lbar, acc, shift):
    """Riccati adjoint, t = 0..T-1 (eigen-shift treated as a constant)."""
    A, B, R = c["Aa"], c["Ba"], c["R"]
    T, b = L.shape[0], A.shape[0]
    u = B.shape[1]
    Sn = np.zeros((b, b))
    for t in range(T):
        S = Sric[t]
        SA, SB = S @ A, S @ B
        H = R + B.T @ SB
        G = B.T @ SA
        Hti = np.linalg.inv(H + shift[t] * np.eye(u))
        Lt = L[t]
        acc["Q"] += Sn
        acc["Aa"] += 2.0 * SA @ Sn
        Sb = A @ Sn @ A.T
        Lb = Lbar[t] + 2.0 * H @ Lt @ Sn + 2.0 * G @ Sn
        Hb = Lt @ Sn @ Lt.T
        Gb = 2.0 * Lt @ Sn
        Gb = Gb - Hti.T @ Lb
        Hb = Hb - Hti.T @ Lb @ Lt.T
        Hb = sym(Hb)
        acc["R"] += Hb
        acc["Ba"] += 2.0 * SB @ Hb + SA @ Gb.T
        acc["Aa"] += SB @ Gb
        Sb = Sb + B @ Hb @ B.T + sym(B @ Gb @ A.T)
        Sn = sym(Sb)
    acc["Q"] += Sn      # Qf = Q[-1] (lqg/utils.py:30)
    return Sn


def derived_to_base(act, dyn, acc):
    """Chain the derived-constant cotangents back to the 12 base matrices."""
    g_act = {k: np.zeros_like(act[k]) for k in ("A", "B", "F", "V", "W", "Q", "R")}
    g_dyn = {k: np.zeros_like(dyn[k]) for k in ("A", "B", "F", "V", "W")}
    g_act["A"] += acc["Aa"]; g_act["B"] += acc["Ba"]; g_act["F"] += acc["Fa"]
    g_act["Q"] += sym(acc["Q"]); g_act["R"] += sym(acc["R"])
    g_act["V"] += 2.0 * sym(acc["VVa"]) @ act["V"]
    g_act["W"] += 2.0 * sym(acc["WWa"]) @ act["W"]
    g_dyn["A"] += acc["Ad"]; g_dyn["B"] += acc["Bd"]
    # FAd = Fd Ad ; FAa = Fa Aa ; D = Fd Bd - Fa Ba
    g_dyn["F"] += acc["FAd"] @ dyn["A"].T + acc["D"] @ dyn["B"].T
    g_dyn["A"] += dyn["F"].T @ acc["FAd"]
    g_dyn["B"] += dyn["F"].T @ acc["D"]
    g_act["F"] += acc["FAa"] @ act["A"].T - acc["D"] @ act["B"].T
    g_act["A"] += act["F"].T @ acc["FAa"]
    g_act["B"] -= act["F"].T @ acc["D"]
    # Om = FN Fd^T + Wd Wd^T ; FN = Fd N11 ; N11 = Vd Vd^T   (Om, N11 symmetric)
    Omb = sym(acc["Om"])
    N11 = dyn["V"] @ dyn["V"].T
    FNb = acc["FN"] + Omb @ dyn["F"]
    g_dyn["F"] += Omb @ (dyn["F"] @ N11) + FNb @ N11
    g_dyn["W"] += 2.0 * Omb @ dyn["W"]
    N11b = sym(acc["N11"] + dyn["F"].T @ FNb)
    g_dyn["V"] += 2.0 * N11b @ dyn["V"]
    return g_act, g_dyn


def value_and_grad(act, dyn, X, w=None, Sigma0=None):
    """ll[N] and d(sum_i w_i ll_i)/d(base matrices).  Returns ll, (g_act, g_dyn)."""
    X = np.asarray(X, dtype=np.float64)
    N, T1, d = X.shape
    w = np.ones(N) if w is None else np.asarray(w, dtype=np.float64)
    ll, f = forward(act, dyn, X, Sigma0=Sigma0)
    c = f["c"]
    acc = {k: np.zeros_like(v) for k, v in c.items()}
    sums = trial_rev(f["rec"], X, f["hist"], w)
    Lbar, Kbar = cov_rev(c, f["L"], f["K"], f["Cs"], sums, w.sum(), d, acc)
    kf_rev(c, f["K"], f["Pkf"], Kbar, acc, own_sigma0=Sigma0 is None)
    lqr_rev(c, f["L"], f["Sric"], Lbar, acc, f["shift"])
    return ll, derived_to_base(act, dyn, acc)

"""Debug helper: stage-by-stage check of the large-system (c4) CUDA path against the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lqg_b200 import abi
from tests import helpers as H
from oracle import lqg_np as O

dev = torch.device("cuda:0")
lib = abi.load_library()
case = H.Case("pmdelay2", S=3, T=40, N=50, d=2, want_grad=True)
dims = case.lqgk_dims()
act, dyn = case.tensors(dev, torch.float64)
ws = H.workspace(lib, dims, abi.MODE_GAINS, dev)
L, l, Hh = lib.lqr_backward(dims, act, ws=ws, stream=H.stream_of(dev))
K = lib.kf_forward(dims, act, ws=ws, stream=H.stream_of(dev))
torch.cuda.synchronize()
sa, _ = O.make_system(case.mats[0], case.T)
Lo, lo, Ho = O.lqr_backward(sa)
Ko = O.kf_forward(sa, sa["V"][0] @ sa["V"][0].T)
print("L err", np.abs(L[0].cpu().numpy() - Lo).max(), "K err", np.abs(K[0].cpu().numpy() - Ko).max(), "nan L", torch.isnan(L).any().item(), "nan K", torch.isnan(K).any().item())
x_tm = lib.pack_obs(torch.tensor(case.X, device=dev), stream=H.stream_of(dev))
ws = H.workspace(lib, dims, abi.MODE_FWD, dev)
ll = lib.loglik_fwd(dims, act, dyn, x_tm, ws=ws, stream=H.stream_of(dev))
torch.cuda.synchronize()
print("ll fwd", ll[0, :4].cpu().numpy(), "oracle", case.ll[0, :4], "launches", lib.last_launch_count())
ws = H.workspace(lib, dims, abi.MODE_VJP, dev)
ll, oa, od, _ = lib.loglik_vjp(dims, act, dyn, x_tm, ws=ws, stream=H.stream_of(dev))
torch.cuda.synchronize()
print("ll vjp", ll[0, :4].cpu().numpy())
for k in abi.ACTOR_KEYS:
    g = oa[k][0].cpu().numpy(); r = case.ga[0][k]
    print("actor", k, "err", np.abs(g - r).max(), "scale", np.abs(r).max(), "nan", np.isnan(g).any())
for k in abi.DYN_KEYS:
    g = od[k][0].cpu().numpy(); r = case.gd[0][k]
    print("dyn", k, "err", np.abs(g - r).max(), "scale", np.abs(r).max(), "nan", np.isnan(g).any())

# ---- raw workspace dump (FWD-mode plan: cst | L | K | rec | ll | scr, 256-byte aligned)
from oracle import adjoint_np as AD
def tri(n): return n * (n + 1) // 2
x, b, u, y, d = case.dims
total = b*b + b*u + y*b + tri(b) + tri(u) + tri(b) + tri(b) + tri(y) + tri(b) + x*x + x*u + y*x + y*b + y*u + tri(x) + y*x + tri(y) + b + u + u*b + b
Sc, T = 32, case.T
up = lambda v: (v + 255) // 256 * 256
n = x + b; r = n - d
REC = (n*n + r*d + tri(d) + 1 + 3) // 4 * 4
off = 0
o_cst = off; off = up(off + 8*total*Sc)
o_L = off; off = up(off + 8*T*u*b*Sc)
o_K = off; off = up(off + 8*T*b*y*Sc)
o_rec = off; off = up(off + 4*Sc*T*REC)
ws = H.workspace(lib, dims, abi.MODE_FWD, dev)
ws.zero_()
ll = lib.loglik_fwd(dims, act, dyn, x_tm, ws=ws, stream=H.stream_of(dev))
torch.cuda.synchronize()
rec = ws[o_rec:o_rec + 4*Sc*T*REC].view(torch.float32).view(Sc, T, REC).cpu().numpy()
c = AD.derive_consts(case.mats[0][0], case.mats[0][1])
Lo_, Sric, shift = AD.lqr_fwd(c, T)
Ko_, Pkf = AD.kf_fwd(c, T)[:2]
Cs, ro, J0 = AD.cov_fwd(c, Lo_, Ko_, d)
for t in (0, 1, T - 1):
    Fg = rec[0, t, :n*n].reshape(n, n).copy(); Fg[:d] *= -1
    print("t", t, "F err", np.abs(Fg - ro["F"][t]).max(), "J err", np.abs(rec[0, t, n*n:n*n + r*d].reshape(r, d) - ro["J"][t]).max(),
          "logdet", rec[0, t, n*n + r*d + tri(d)], ro["logdet"][t], "Linv", rec[0, t, n*n + r*d:n*n + r*d + tri(d)], ro["Linv"][t].ravel())
print("nan in rec:", np.isnan(rec[:3]).sum(), "inf:", np.isinf(rec[:3]).sum())

# ---- replay the per-trial forward recursion in float32 NumPy from the GPU's own records
def replay(rec_s, X, dtype):
    N, T1, d_ = X.shape
    c_ = np.zeros((N, r), dtype); ll_ = np.zeros(N)
    for t in range(T1 - 1):
        R_ = rec_s[t].astype(dtype)
        Fm = R_[:n*n].reshape(n, n); J = R_[n*n:n*n + r*d].reshape(r, d)
        Li = np.zeros((d, d), dtype); Li[np.tril_indices(d)] = R_[n*n + r*d:n*n + r*d + tri(d)]
        x0, x1 = X[:, t].astype(dtype), X[:, t + 1].astype(dtype)
        e = x1 + x0 @ Fm[:d, :d].T + c_ @ Fm[:d, d:].T
        z = e @ Li.T
        ll_ += -0.5 * (z * z).sum(1) - R_[n*n + r*d + tri(d)] - 0.5 * d * np.log(2 * np.pi)
        c_ = x0 @ Fm[d:, :d].T + c_ @ Fm[d:, d:].T + e @ J.T
    return ll_, np.abs(c_).max()
print("replay f32:", replay(rec[0], case.X, np.float32)[0][:4], "max|c|", replay(rec[0], case.X, np.float32)[1])
print("replay f64:", replay(rec[0], case.X, np.float64)[0][:4])
print("J scale per step:", [float(np.abs(ro["J"][t]).max()) for t in (0, 1, 20, 39)])
np.set_printoptions(linewidth=200, precision=4, suppress=False)
print("GPU J t=0:", rec[0, 0, n*n:n*n + r*d])
print("ORC J t=0:", ro["J"][0].ravel())
print("GPU J t=20:", rec[0, 20, n*n:n*n + r*d])
print("ORC J t=20:", ro["J"][20].ravel())

# ---- constants block as packed on the GPU vs the oracle's derived constants
cst = ws[o_cst:o_cst + 8*total*Sc].view(torch.float64).view(total, Sc).cpu().numpy()[:, 0]
sizes = [("Aa", b*b), ("Ba", b*u), ("Fa", y*b), ("Q", tri(b)), ("R", tri(u)), ("Qf", tri(b)), ("VVa", tri(b)), ("WWa", tri(y)), ("Sig0", tri(b)),
         ("Ad", x*x), ("Bd", x*u), ("FAd", y*x), ("FAa", y*b), ("Dm", y*u), ("N11", tri(x)), ("FN", y*x), ("Om", tri(y))]
o = 0; blk = {}
for k_, n_ in sizes:
    blk[k_] = cst[o:o + n_]; o += n_
def packed(M):
    return np.array([M[i, j] for i in range(M.shape[0]) for j in range(i + 1)])
for k_, ref in (("Aa", c["Aa"].ravel()), ("Ad", c["Ad"].ravel()), ("FAd", c["FAd"].ravel()), ("FAa", c["FAa"].ravel()), ("Dm", c["D"].ravel()),
                ("N11", packed(c["N11"])), ("FN", c["FN"].ravel()), ("Om", packed(c["Om"]))):
    print("const", k_, "err", np.abs(blk[k_] - ref).max(), "scale", np.abs(ref).max())
Kg = ws[o_K:o_K + 8*T*b*y*Sc].view(torch.float64).view(T, b*y, Sc).cpu().numpy()[:, :, 0].reshape(T, b, y)
Lg = ws[o_L:o_L + 8*T*u*b*Sc].view(torch.float64).view(T, u*b, Sc).cpu().numpy()[:, :, 0].reshape(T, u, b)
print("K ws err", np.abs(Kg - Ko_).max(), "L ws err", np.abs(Lg - Lo_).max())
N0 = AD.joint_N(c, Kg[0]); print("Sig0 uo (oracle consts, GPU K):", N0[d:d+6, :d].ravel())

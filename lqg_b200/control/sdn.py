"""Signal-dependent-noise LQG gains (Todorov 2005) -- an EXTENSION of the reference's ``lqg.control`` (the reference has no
signal-dependent noise: docs/README.md:60-62).  ``solve`` runs the alternating control / estimator iterations in one CUDA
kernel (``k_sdn_gains``, csrc/lqgk_sdn.cuh); mathematical spec and validation: ``oracle/sdn_np.py``, ``tests/test_sdn.py``.

Convention (the paper's predictor form, not the reference's filter form): ``u_t = -L_t xhat_t``,
``xhat_{t+1} = A xhat_t + B u_t + K_t (y_t - H xhat_t)``.  Without multiplicative noise one sweep gives
``L = -lqr.backward(...).L``."""
from typing import NamedTuple, Optional, Sequence

import torch

from lqg_b200 import abi


class SDNGains(NamedTuple):
    L: torch.Tensor      # [S, T, u, b]
    K: torch.Tensor      # [S, T, b, y]
    cost: torch.Tensor   # [S] expected total cost


def solve(A, B, H, Q, R, Om_xi, Om_omega, Sigma1, xhat1, T: int, C: Optional[Sequence] = None, D: Optional[Sequence] = None,
          Qf=None, sweeps: int = 10, form: str = "predictor") -> SDNGains:
    """All arguments CUDA tensors (cast to float64); any of them may carry a leading parameter-sample axis S.
    ``C``: control-dependent noise matrices [nc, b, u] (or [S, nc, b, u]); ``D``: state-dependent observation noise [nd, y, b].
    ``form``: "predictor" = the paper's convention (u = -L xhat, xhat' = A xhat + B u + K (y - H xhat)); "filter" = the
    reference's (u = L xhat, xhat' = xp + K (y' - H xp), lqg/system.py:110-124): without multiplicative noise exactly
    lqr.backward / kf.forward, and the gains plug into ``System.log_likelihood_sdn(gains=(L, K))``."""
    dev = A.device
    if dev.type != "cuda":
        raise RuntimeError("lqg_b200 has no CPU fallback: tensors must live on a CUDA device")
    f = lambda v: None if v is None else torch.as_tensor(v, device=dev).to(torch.float64)
    mats = dict(A=f(A), B=f(B), H=f(H), Q=f(Q), R=f(R), Qf=f(Qf), Om_xi=f(Om_xi), Om_omega=f(Om_omega), Sigma1=f(Sigma1), xhat1=f(xhat1),
                C=None if C is None else f(C if torch.is_tensor(C) else torch.stack(list(C))),
                D=None if D is None else f(D if torch.is_tensor(D) else torch.stack(list(D))))
    if form not in ("predictor", "filter"):
        raise ValueError(form)
    L, K, cost = abi.load_library().sdn_gains(mats, T, sweeps, stream=torch.cuda.current_stream(dev).cuda_stream,
                                              filter_form=form == "filter")
    return SDNGains(L, K, cost)


def solve_for_actor(system, signal_dep_noise=None, obs_dep_noise=None, C=None, D=None, Sigma0=None, xhat0=None, sweeps: int = 6) -> SDNGains:
    """Filter-form gains of ``system``'s ACTOR model under signal-dependent noise (what replaces ``lqr.backward(actor)`` /
    ``kf.forward(actor)`` of lqg/system.py:157-161 when the actor plans and filters knowing about the multiplicative noise).
    Noise either as per-channel scales on the actor's own B / F (``channel_noise`` on the actor spec) or as explicit
    ``C[(S,) nc, b, u]``, ``D[(S,) nd, y, b]``.  Time-invariant specs."""
    a = system.actor
    base = lambda M: M[..., 0, :, :]
    A, Bm, F, V, W, Q, R = (base(getattr(a, k)) for k in ("A", "B", "F", "V", "W", "Q", "R"))
    if C is None and signal_dep_noise is not None:
        C = _channel(Bm, signal_dep_noise, control=True)
    if D is None and obs_dep_noise is not None:
        D = _channel(F, obs_dep_noise, control=False)
    mT = lambda M: M.transpose(-1, -2)
    Om_xi, Om_om = V @ mT(V), W @ mT(W)
    S0 = Om_xi if Sigma0 is None else torch.as_tensor(Sigma0, dtype=A.dtype, device=A.device)     # system.py:158-161 default
    xh = torch.zeros(A.shape[-1], dtype=A.dtype, device=A.device) if xhat0 is None else torch.as_tensor(xhat0, dtype=A.dtype, device=A.device)
    Qf = a.Qf if a.Qf.dim() == Q.dim() else a.Qf[..., 0, :, :] if a.Qf.dim() > Q.dim() else a.Qf
    return solve(A, Bm, F, Q, R, Om_xi, Om_om, S0, xh, a.A.shape[-3], C=C, D=D, Qf=Qf, sweeps=sweeps, form="filter")


# ------------------------------------------------------------------------------------------------ likelihood side
def _channel(M: torch.Tensor, scale, control: bool) -> torch.Tensor:
    """control: [B[:, i] e_i^T]_i from the input map M = B[(S,) x, u]; else [e_j e_j^T F]_j from the observation map M = F[(S,) y, x]."""
    scale = torch.as_tensor(scale, dtype=M.dtype, device=M.device)
    if control:
        u = M.shape[-1]
        mats = [M * torch.nn.functional.one_hot(torch.tensor(i), u).to(M) for i in range(u)]
    else:
        y = M.shape[-2]
        mats = [M * torch.nn.functional.one_hot(torch.tensor(j), y).to(M)[:, None] for j in range(y)]
    out = torch.stack(mats, -3)
    if scale.dim() == 1 and out.dim() == 3:
        out = out.unsqueeze(0)
    return out * scale.reshape(scale.shape + (1, 1, 1)) if scale.dim() else out * scale


def channel_noise(system, scale, kind: str) -> torch.Tensor:
    """Per-channel proportional noise matrices for ``System.log_likelihood_sdn`` in the reference's filter-form model
    (lqg/system.py:110-124 plus the multiplicative terms, oracle/sdn_np.py):

    * ``kind="control"``: ``C_i = scale * B_d[:, i] e_i^T`` (one per control dimension): the process noise injected through
      control channel i is proportional to ``u_i`` -- the motor noise of Todorov (2005) along the plant's own input map;
    * ``kind="observation"``: ``D_j = scale * e_j e_j^T F_d`` (one per observation channel): the noise of observation j is
      proportional to the observed signal ``(F_d x)_j`` itself.

    ``scale``: scalar or batched like the model parameters ([S]).  Returns [(S,) nc, x, u] / [(S,) nd, y, x]."""
    dyn = system.dynamics
    if kind == "control":
        return _channel(dyn.B[..., 0, :, :], scale, control=True)
    if kind == "observation":
        return _channel(dyn.F[..., 0, :, :], scale, control=False)
    raise ValueError(kind)


def value_and_grad_fd(fn, theta: torch.Tensor, rel_step: float = 1e-5, abs_step: float = 1e-7):
    """Value and gradient of ``fn`` at ``theta[P]`` by central differences evaluated as ONE batched call:
    ``fn(Theta[2P+1, P]) -> [2P+1]`` (e.g. the summed log-likelihood of a model built from batched parameters).  The
    signal-dependent-noise likelihood has no adjoint kernel; its evaluations are latency-bound and the parameter-sample axis
    of the kernels is free, so the 2P+1 evaluations cost about as much as one.  FP64 recommended."""
    theta = theta.detach()
    P = theta.numel()
    h = (theta.abs() * rel_step).clamp_min(abs_step)
    Theta = theta.repeat(2 * P + 1, 1)
    idx = torch.arange(P, device=theta.device)
    Theta[1 + idx, idx] += h
    Theta[1 + P + idx, idx] -= h
    f = fn(Theta)
    return f[0], (f[1:1 + P] - f[1 + P:]) / (2 * h)

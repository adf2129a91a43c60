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
          Qf=None, sweeps: int = 10) -> SDNGains:
    """All arguments CUDA tensors (cast to float64); any of them may carry a leading parameter-sample axis S.
    ``C``: control-dependent noise matrices [nc, b, u] (or [S, nc, b, u]); ``D``: state-dependent observation noise [nd, y, b]."""
    dev = A.device
    if dev.type != "cuda":
        raise RuntimeError("lqg_b200 has no CPU fallback: tensors must live on a CUDA device")
    f = lambda v: None if v is None else torch.as_tensor(v, device=dev).to(torch.float64)
    mats = dict(A=f(A), B=f(B), H=f(H), Q=f(Q), R=f(R), Qf=f(Qf), Om_xi=f(Om_xi), Om_omega=f(Om_omega), Sigma1=f(Sigma1), xhat1=f(xhat1),
                C=None if C is None else f(C if torch.is_tensor(C) else torch.stack(list(C))),
                D=None if D is None else f(D if torch.is_tensor(D) else torch.stack(list(D))))
    L, K, cost = abi.load_library().sdn_gains(mats, T, sweeps, stream=torch.cuda.current_stream(dev).cuda_stream)
    return SDNGains(L, K, cost)


# ------------------------------------------------------------------------------------------------ likelihood side
def channel_noise(system, scale, kind: str) -> torch.Tensor:
    """Per-channel proportional noise matrices for ``System.log_likelihood_sdn`` in the reference's filter-form model
    (lqg/system.py:110-124 plus the multiplicative terms, oracle/sdn_np.py):

    * ``kind="control"``: ``C_i = scale * B_d[:, i] e_i^T`` (one per control dimension): the process noise injected through
      control channel i is proportional to ``u_i`` -- the motor noise of Todorov (2005) along the plant's own input map;
    * ``kind="observation"``: ``D_j = scale * e_j e_j^T F_d`` (one per observation channel): the noise of observation j is
      proportional to the observed signal ``(F_d x)_j`` itself.

    ``scale``: scalar or batched like the model parameters ([S]).  Returns [(S,) nc, x, u] / [(S,) nd, y, x]."""
    dyn = system.dynamics
    Bd, Fd = dyn.B[..., 0, :, :], dyn.F[..., 0, :, :]
    scale = torch.as_tensor(scale, dtype=Bd.dtype, device=Bd.device)
    if kind == "control":
        u = Bd.shape[-1]
        mats = [Bd * torch.nn.functional.one_hot(torch.tensor(i), u).to(Bd) for i in range(u)]            # B_d[:, i] e_i^T
    elif kind == "observation":
        y = Fd.shape[-2]
        mats = [Fd * torch.nn.functional.one_hot(torch.tensor(j), y).to(Fd)[:, None] for j in range(y)]   # e_j e_j^T F_d
    else:
        raise ValueError(kind)
    M = torch.stack(mats, -3)
    if scale.dim() == 1 and M.dim() == 3:
        M = M.unsqueeze(0)
    return M * scale.reshape(scale.shape + (1, 1, 1)) if scale.dim() else M * scale


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

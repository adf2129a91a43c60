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

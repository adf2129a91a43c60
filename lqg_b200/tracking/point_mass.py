"""``PointMassBoundedActor`` -- same construction as ``lqg/tracking/point_mass.py:7-144``: point mass with viscous
damping and first-order muscle activation, zero-order-hold discretisation via the matrix exponential, process
noise from the top-right block of the Van-Loan exponential, made PSD by eigenvalue clipping and factored with an
UPPER Cholesky factor (the reference's ``jax.scipy.linalg.cholesky`` default, point_mass.py:123-125)."""
import torch

from lqg_b200.system import Actor, System
from lqg_b200.tracking import _build as B


def discretize_linear_system(A, Bm, dt):
    """point_mass.py:50-79 (zero-order hold)."""
    n, m = A.shape[-1], Bm.shape[-1]
    M = torch.zeros(A.shape[:-2] + (n + m, n + m), dtype=A.dtype, device=A.device)
    M[..., :n, :n] = A
    M[..., :n, n:] = Bm
    E = torch.linalg.matrix_exp(M * dt)
    return E[..., :n, :n], E[..., :n, n:]


def van_loan_discretization(A, G, dt, Qc=None):
    """point_mass.py:82-110 (returns the top-right block, as the reference does)."""
    n = A.shape[-1]
    if Qc is None:
        Qc = torch.eye(G.shape[-1], dtype=A.dtype, device=A.device)
    Q = G @ Qc @ G.transpose(-1, -2)
    M = torch.cat([torch.cat([A, Q], -1), torch.cat([torch.zeros_like(A), -A.transpose(-1, -2)], -1)], -2)
    return torch.linalg.matrix_exp(M * dt)[..., :n, n:]


def make_psd(M, eps=1e-6):
    """point_mass.py:128-144."""
    w, U = torch.linalg.eigh(0.5 * (M + M.transpose(-1, -2)))
    return U @ torch.diag_embed(torch.clamp(w, min=eps)) @ U.transpose(-1, -2)


def point_mass_dynamics_matrices(damping, m, tau, action_variability, dt):
    """point_mass.py:113-125."""
    z, o = torch.zeros_like(damping), torch.ones_like(damping)
    A_c = torch.stack([torch.stack([z, o, z], -1), torch.stack([z, -damping / m, 1.0 / m], -1),
                       torch.stack([z, z, -1.0 / tau], -1)], -2)
    B_c = torch.stack([z, z, 1.0 / tau], -1).unsqueeze(-1)
    A, Bm = discretize_linear_system(A_c, B_c, dt)
    Qd = make_psd(van_loan_discretization(A_c, 1e-2 * action_variability[..., None, None] * B_c, dt))
    V = torch.linalg.cholesky(Qd, upper=True)
    return A, Bm, V


class PointMassBoundedActor(System):
    def __init__(self, process_noise=1.0, action_variability=1e-3, sigma_target=6.0, sigma_cursor=6.0, action_cost=0.01,
                 dt=1.0 / 60.0, T=1000, damping=0.1, m=1.0, tau=0.0015, dtype=None, device=None):
        (pn, av, st, sc, ac, dm, ms, ta), batch, dtype, device = B.canon(
            [process_noise, action_variability, sigma_target, sigma_cursor, action_cost, damping, m, tau], dtype, device)
        # expm / eigh / cholesky of the tiny continuous-time system in float64 for stability, then cast
        A3, B3, V3 = point_mass_dynamics_matrices(dm.double(), ms.double(), ta.double(), av.double(), dt)
        A3, B3, V3 = A3.to(dtype), B3.to(dtype), V3.to(dtype)
        A = torch.zeros(batch + (4, 4), dtype=dtype, device=device)
        A[..., 0, 0] = 1.0
        A[..., 1:, 1:] = A3
        Bm = torch.cat([torch.zeros(batch + (1, 1), dtype=dtype, device=device), B3], -2)
        V = torch.zeros(batch + (4, 4), dtype=dtype, device=device)
        V[..., 0, 0] = pn
        V[..., 1:, 1:] = V3
        F = torch.eye(3, 4, dtype=dtype, device=device)
        W = B.diag([st, sc, sc])
        Q = torch.zeros(4, 4, dtype=dtype, device=device)
        Q[:2, :2] = B.const([[1.0, -1.0], [-1.0, 1.0]], dtype, device)
        R = torch.eye(1, dtype=dtype, device=device) * (ac * dt)[..., None, None]
        spec = Actor(A=A, B=Bm, F=F, V=V, W=W, Q=Q, R=R, T=T)
        super().__init__(actor=spec, dynamics=spec)

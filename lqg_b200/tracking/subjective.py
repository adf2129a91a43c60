"""``SubjectiveActor`` -- same constructor as ``lqg/tracking/subjective.py:15-47`` (actor with a subjective
velocity state per axis; dynamics as in BoundedActor); parameters may be batched tensors."""
import torch

from lqg_b200.system import Actor, Dynamics, System
from lqg_b200.tracking import _build as B
from lqg_b200.tracking._build import swap_dims  # noqa: F401  (re-export, reference name)


class SubjectiveActor(System):
    def __init__(self, dim=1, process_noise=1.0, action_cost=1.0, action_variability=0.5, subj_noise=1.0,
                 subj_vel_noise=0.5, sigma_target=6.0, sigma_cursor=6.0, dt=1.0 / 60, T=1000, dtype=None, device=None):
        self.dim = dim
        (pn, ac, av, sn, svn, st, sc), batch, dtype, device = B.canon(
            [process_noise, action_cost, action_variability, subj_noise, subj_vel_noise, sigma_target, sigma_cursor],
            dtype, device)
        A = torch.eye(2 * dim, dtype=dtype, device=device)
        Bm = B.block_diag_const([[0.0], [1.0 * dt]], dim, dtype, device)
        F = torch.eye(2 * dim, dtype=dtype, device=device)
        V = B.diag([pn, av] * dim)
        W = B.diag([st, sc] * dim)
        dyn = Dynamics(A=A, B=Bm, F=F, V=V, W=W, T=T)

        A = B.block_diag_const([[1.0, 0.0, dt], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]], dim, dtype, device)
        Bm = B.block_diag_const([[0.0], [1.0 * dt], [0.0]], dim, dtype, device)
        F = B.block_diag_const([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], dim, dtype, device)
        V = B.diag([sn, av, svn] * dim)
        Q = B.block_diag_const([[1.0, -1.0, 0.0], [-1.0, 1.0, 0.0], [0.0, 0.0, 0.0]], dim, dtype, device)
        R = torch.eye(dim, dtype=dtype, device=device) * ac[..., None, None]

        dims = B.index(swap_dims(A.shape[0], dim), device)     # cached device index (no host-to-device copy per construction)
        A = A.index_select(0, dims).index_select(1, dims)
        Bm = Bm.index_select(0, dims)
        V = V.index_select(-2, dims)
        F = F.index_select(1, dims)
        Q = Q.index_select(0, dims).index_select(1, dims)
        act = Actor(A=A, B=Bm, F=F, V=V, W=W, Q=Q, R=R, T=T)
        super().__init__(actor=act, dynamics=dyn)
        if dim > 1:   # identical independent axes: the likelihood factorises (System.log_likelihood)
            self._axis_system = SubjectiveActor(dim=1, process_noise=process_noise, action_cost=action_cost,
                                                action_variability=action_variability, subj_noise=subj_noise,
                                                subj_vel_noise=subj_vel_noise, sigma_target=sigma_target,
                                                sigma_cursor=sigma_cursor, dt=dt, T=T, dtype=dtype, device=device)

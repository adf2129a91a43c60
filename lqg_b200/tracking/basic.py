"""``TrackingTask`` / ``BoundedActor`` / ``OptimalActor`` / ``RelativeObservationBoundedActor`` -- same constructors
as ``lqg/tracking/basic.py:7-124``.  Parameters may be Python floats or torch tensors; tensors of shape ``(S,)`` give a
batched system (one parameter sample per row), which is how ``jax.vmap`` over parameters is expressed here."""
import torch

from lqg_b200.system import Actor, System
from lqg_b200.tracking import _build as B


class TrackingTask(System):
    def __init__(self, dim=1, process_noise=1.0, action_variability=0.5, sigma_target=6.0, sigma_cursor=6.0,
                 action_cost=1.0, dt=1.0 / 60.0, T=1000, dtype=None, device=None):
        self.dim = dim
        self.process_noise = process_noise
        (pn, av, st, sc, ac), batch, dtype, device = B.canon(
            [process_noise, action_variability, sigma_target, sigma_cursor, action_cost], dtype, device)
        d = 2 * dim
        A = torch.eye(d, dtype=dtype, device=device)
        Bm = dt * B.block_diag_const([[0.0], [1.0]], dim, dtype, device)
        F = torch.eye(d, dtype=dtype, device=device)
        V = B.diag([pn, av] * dim)
        W = B.diag([st, sc] * dim)
        Q = B.block_diag_const([[1.0, -1.0], [-1.0, 1.0]], dim, dtype, device)
        R = torch.eye(dim, dtype=dtype, device=device) * ac[..., None, None]
        spec = Actor(A=A, B=Bm, F=F, V=V, W=W, Q=Q, R=R, T=T)
        super().__init__(actor=spec, dynamics=spec)
        if dim > 1:   # identical independent axes: the likelihood factorises (System.log_likelihood)
            self._axis_system = TrackingTask(dim=1, process_noise=process_noise, action_variability=action_variability,
                                             sigma_target=sigma_target, sigma_cursor=sigma_cursor, action_cost=action_cost,
                                             dt=dt, T=T, dtype=dtype, device=device)


class BoundedActor(TrackingTask):
    def __init__(self, dim=1, process_noise=1.0, action_variability=0.5, sigma_target=6.0, sigma_cursor=6.0,
                 action_cost=1.0, dt=1.0 / 60, T=1000, dtype=None, device=None):
        super().__init__(dim=dim, process_noise=process_noise, action_variability=action_variability,
                         sigma_target=sigma_target, sigma_cursor=sigma_cursor, action_cost=action_cost, dt=dt, T=T,
                         dtype=dtype, device=device)


class OptimalActor(TrackingTask):
    def __init__(self, dim=1, process_noise=1.0, action_variability=0.5, sigma_target=6.0, sigma_cursor=6.0,
                 dt=1.0 / 60, T=1000, dtype=None, device=None):
        super().__init__(dim=dim, process_noise=process_noise, action_variability=action_variability,
                         sigma_target=sigma_target, sigma_cursor=sigma_cursor, action_cost=1e-3, dt=dt, T=T,
                         dtype=dtype, device=device)


class RelativeObservationBoundedActor(System):
    def __init__(self, dim=1, process_noise=1.0, action_variability=0.5, sigma=6.0, action_cost=1.0, dt=1.0 / 60.0,
                 T=1000, dtype=None, device=None):
        self.dim = dim
        self.process_noise = process_noise
        (pn, av, sg, ac), batch, dtype, device = B.canon([process_noise, action_variability, sigma, action_cost], dtype, device)
        d = 2 * dim
        A = torch.eye(d, dtype=dtype, device=device)
        Bm = dt * B.block_diag_const([[0.0], [1.0]], dim, dtype, device)
        F = B.block_diag_const([[1.0, -1.0]], dim, dtype, device)
        V = B.diag([pn, av] * dim)
        W = B.diag([sg] * dim)
        Q = B.block_diag_const([[1.0, -1.0], [-1.0, 1.0]], dim, dtype, device)
        R = torch.eye(dim, dtype=dtype, device=device) * ac[..., None, None]
        spec = Actor(A=A, B=Bm, F=F, V=V, W=W, Q=Q, R=R, T=T)
        super().__init__(actor=spec, dynamics=spec)
        if dim > 1:
            self._axis_system = RelativeObservationBoundedActor(dim=1, process_noise=process_noise,
                                                                action_variability=action_variability, sigma=sigma,
                                                                action_cost=action_cost, dt=dt, T=T, dtype=dtype, device=device)

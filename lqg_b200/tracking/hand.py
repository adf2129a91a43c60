"""``HandMotionModelTrackingTask`` -- the hand-dynamics model defined in the reference's ``notebooks/HandModel.ipynb``
(cell 2): state ``[target, position, velocity, force, activation]``, second-order muscle filter with time constant
``tau``, point mass ``m``; only target and cursor position are observed (``F = eye(2, 5)``).

With the notebook's default noise model (process noise only on the target and the activation) the cursor position has no
process noise of its own, so the experimenter-side likelihood is degenerate at the first step exactly as in the
reference; pass ``position_noise > 0`` to regularise it.  Simulation and the gain recursions are always well posed."""
import torch

from lqg_b200.system import Actor, System
from lqg_b200.tracking import _build as B


class HandMotionModelTrackingTask(System):
    def __init__(self, process_noise=1.0, action_variability=0.5, sigma_target=6.0, sigma_cursor=6.0, action_cost=1.0,
                 dt=1.0 / 60.0, m=1.0, tau=0.04, T=1000, position_noise=0.0, dtype=None, device=None):
        self.process_noise = process_noise
        (pn, av, st, sc, ac, ms, ta, pz), batch, dtype, device = B.canon(
            [process_noise, action_variability, sigma_target, sigma_cursor, action_cost, m, tau, position_noise], dtype, device)
        A = torch.zeros(batch + (5, 5), dtype=dtype, device=device)
        A[..., 0, 0] = 1.0
        A[..., 1, 1] = 1.0
        A[..., 1, 2] = dt
        A[..., 2, 2] = 1.0
        A[..., 2, 3] = dt / ms
        A[..., 3, 3] = 1.0 - dt / ta
        A[..., 3, 4] = dt / ta
        A[..., 4, 4] = 1.0 - dt / ta
        Bm = torch.zeros(batch + (5, 1), dtype=dtype, device=device)
        Bm[..., 4, 0] = dt / ta
        F = torch.eye(2, 5, dtype=dtype, device=device)
        z = torch.zeros_like(pn)
        V = B.diag([pn, pz, z, z, av])
        W = B.diag([st, sc])
        Q = torch.zeros(5, 5, dtype=dtype, device=device)
        Q[:2, :2] = B.const([[1.0, -1.0], [-1.0, 1.0]], dtype, device)
        R = torch.eye(1, dtype=dtype, device=device) * ac[..., None, None]
        spec = Actor(A=A, B=Bm, F=F, V=V, W=W, Q=Q, R=R, T=T)
        super().__init__(actor=spec, dynamics=spec)

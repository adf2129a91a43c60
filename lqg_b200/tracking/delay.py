"""``delay_system`` / ``TemporalDelayModel`` / ``DelayedSubjectiveActor`` -- same augmentation as
``lqg/tracking/delay.py:9-51`` (shift register of ``delay`` past states; the observation reads the oldest copy).
The augmented dimension tuples are only runnable through the CUDA path if compiled (csrc/lqgk_dims.h)."""
import torch

from lqg_b200.spec import LQGSpec
from lqg_b200.system import System
from lqg_b200.tracking.subjective import SubjectiveActor
from lqg_b200.utils import time_stack_spec


def _base(M):
    """Base matrix of a time-invariant stacked array (..., T, r, c) -> (..., r, c)."""
    return M[..., 0, :, :]


def delay_system(spec: LQGSpec, delay: int) -> LQGSpec:
    T = spec.A.shape[-3]
    for M in (spec.A, spec.B, spec.F, spec.V, spec.W, spec.Q, spec.R):
        if M.shape[-3] > 1 and M.stride(-3) != 0 and not bool((M[..., :1, :, :] == M).all()):
            raise NotImplementedError("delay_system: time-varying specs are not supported (the reference augments every time "
                                      "step, lqg/tracking/delay.py:9-41; here the time axis is a stride-0 view of one matrix)")
    A, Bm, F, V, W, Q, R = map(_base, (spec.A, spec.B, spec.F, spec.V, spec.W, spec.Q, spec.R))
    d = A.shape[-1]
    kw = dict(dtype=A.dtype, device=A.device)
    batch = A.shape[:-2]
    n = d * (delay + 1)
    A2 = torch.zeros(batch + (n, n), **kw)
    A2[..., :d, :d] = A
    A2 = A2 + torch.diag(torch.ones(d * delay, **kw), diagonal=-d)
    B2 = torch.cat([Bm] + [torch.zeros_like(Bm)] * delay, -2)
    F2 = torch.cat([torch.zeros(F.shape[:-1] + (F.shape[-1] * delay,), **kw), F], -1)
    V2 = torch.zeros(V.shape[:-2] + (n, n - d + V.shape[-1]), **kw)      # noise enters the current state only
    V2[..., :d, :V.shape[-1]] = V
    Q2 = torch.zeros(Q.shape[:-2] + (n, n), **kw)
    Q2[..., :d, :d] = Q
    return time_stack_spec(A=A2, B=B2, F=F2, V=V2, W=W, Q=Q2, R=R, T=T)


class TemporalDelayModel(System):
    def __init__(self, system, delay):
        dyn = delay_system(system.dynamics, delay=delay)
        act = dyn if system.actor is system.dynamics else delay_system(system.actor, delay=delay)
        super().__init__(actor=act, dynamics=dyn)


class DelayedSubjectiveActor(TemporalDelayModel):
    def __init__(self, process_noise=1.0, c=0.5, action_variability=0.5, subj_noise=1.0, subj_vel_noise=10.0,
                 sigma_target=6.0, sigma_cursor=3.0, dt=1.0 / 60, dtype=None, device=None):
        system = SubjectiveActor(process_noise=process_noise, action_cost=c, action_variability=action_variability,
                                 subj_noise=subj_noise, subj_vel_noise=subj_vel_noise, sigma_target=sigma_target,
                                 sigma_cursor=sigma_cursor, dt=dt, dtype=dtype, device=device)
        super().__init__(system=system, delay=12)

"""Maximum likelihood by Adam on the unconstrained (log) parameters -- counterpart of ``lqg/infer/mle.py:14-25``
(SVI with an empty guide = plain Adam on -log p(x | theta); positive constraint = exp transform as numpyro's
``constraints.positive``).  Every step is one fused CUDA forward+adjoint evaluation."""
import torch

from lqg_b200.infer.models import get_model_params, log_likelihood
from lqg_b200.tracking import BoundedActor


def max_likelihood(x, model=BoundedActor, process_noise=1.0, dt=1.0 / 60, steps=2_000, step_size=0.01, dim=None, **fixed):
    names = [k for k in get_model_params(model) if k not in fixed]
    init = get_model_params(model)
    z = {k: torch.tensor(float(init[k]), device=x.device).log().requires_grad_() for k in names}
    opt = torch.optim.Adam(list(z.values()), lr=step_size)
    losses = []
    for _ in range(steps):
        opt.zero_grad(set_to_none=True)
        theta = {k: v.exp() for k, v in z.items()}
        loss = -log_likelihood(theta, x, model, process_noise=process_noise, dt=dt, dim=dim, **fixed)
        loss.backward()
        opt.step()
        losses.append(loss.detach())
    return {k: v.detach().exp() for k, v in z.items()}, torch.stack(losses)

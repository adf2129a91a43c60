"""Posterior sampling -- counterpart of ``lqg/infer/utils.py:14-41`` (``infer``).

The reference runs numpyro's NUTS; here many chains of plain HMC run in lock-step on the GPU: the chains are the
parameter-sample axis of the kernels, so ONE fused CUDA forward+adjoint call evaluates the potential and its gradient for
all chains (config c5 of BASELINE.json: thousands of vectorised chains).  Step size is adapted during warm-up to a target
acceptance rate; positions live in log space (exp transform + log-Jacobian, as numpyro does for positive parameters)."""
import math

import torch

from lqg_b200.infer.models import get_model_params, log_joint
from lqg_b200.tracking import BoundedActor


class Samples:
    def __init__(self, samples, accept_rate, step_size):
        self._samples, self.accept_rate, self.step_size = samples, accept_rate, step_size

    def get_samples(self, group_by_chain=False):
        if group_by_chain:
            return self._samples
        return {k: v.reshape(-1) for k, v in self._samples.items()}


def infer(x, num_samples, num_warmup, model=BoundedActor, process_noise=1.0, dt=1.0 / 60, method="hmc", num_chains=1, seed=0,
          num_leapfrog=8, step_size=0.02, target_accept=0.8, dim=None, priors=None, **fixed):
    if method not in ("hmc", "nuts"):
        raise ValueError("Please specify a valid inference method (hmc).")
    dev = x.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    names = [k for k in get_model_params(model) if k not in fixed]
    init = get_model_params(model)
    z = torch.stack([torch.full((num_chains,), math.log(float(init[k])), device=dev) for k in names], 1)
    z = z + 0.05 * torch.randn(z.shape, device=dev, generator=gen)

    def potential(zz):
        zz = zz.detach().requires_grad_()
        theta = {k: zz[:, i].exp() for i, k in enumerate(names)}
        lp = log_joint(theta, x, model, priors=priors, process_noise=process_noise, dt=dt, dim=dim, **fixed) + zz.sum(1)
        (g,) = torch.autograd.grad(lp.sum(), zz)
        return -lp.detach(), -g

    U, gU = potential(z)
    eps = torch.full((num_chains, 1), step_size, device=dev)
    out, acc_hist = [], []
    for it in range(num_warmup + num_samples):
        p = torch.randn(z.shape, device=dev, generator=gen)
        H0 = U + 0.5 * (p * p).sum(1)
        zn, pn, Un, gn = z, p - 0.5 * eps * gU, U, gU
        for l in range(num_leapfrog):
            zn = zn + eps * pn
            Un, gn = potential(zn)
            pn = pn - (eps if l < num_leapfrog - 1 else 0.5 * eps) * gn
        H1 = Un + 0.5 * (pn * pn).sum(1)
        a = torch.exp(torch.clamp(H0 - H1, max=0.0))
        a = torch.where(torch.isfinite(a), a, torch.zeros_like(a))
        take = torch.rand(num_chains, device=dev, generator=gen) < a
        z = torch.where(take[:, None], zn, z)
        U = torch.where(take, Un, U)
        gU = torch.where(take[:, None], gn, gU)
        if it < num_warmup:   # Robbins-Monro step-size adaptation per chain
            eps = eps * torch.exp(0.1 * (a[:, None] - target_accept) / math.sqrt(1 + it / 10.0))
        else:
            out.append(z.exp())
            acc_hist.append(a)
    s = torch.stack(out, 1)                                    # (chains, samples, params)
    return Samples({k: s[:, :, i] for i, k in enumerate(names)}, torch.stack(acc_hist, 1).mean().item(), eps.mean().item())

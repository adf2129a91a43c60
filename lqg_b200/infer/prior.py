"""Default priors -- same families and hyper-parameters as ``lqg/infer/prior.py:7-22``, as torch distributions."""
import math

import torch
from torch import distributions as dist


def _t(v):
    return torch.tensor(float(v))


default_prior = {
    "action_cost": dist.LogNormal(_t(-2.0), _t(1.0)),
    "sigma_target": dist.HalfNormal(_t(50.0)),
    "action_variability": dist.HalfNormal(_t(1.0)),
    "signal_dep_noise": dist.HalfNormal(_t(1.0)),
    "sigma_cursor": dist.HalfNormal(_t(12.5)),
    "sigma": dist.HalfNormal(_t(50.0)),
    "subj_noise": dist.HalfNormal(_t(1.0)),
    "subj_vel_noise": dist.HalfNormal(_t(2.0)),
    **{f"sigma_target_{i}": dist.HalfNormal(_t(50.0)) for i in range(6)},
}


def prior():
    return default_prior


def lognormal_params(mu, sigma):
    """``lqg/infer/prior.py:29-30``."""
    return math.log(mu ** 2 / math.sqrt(mu ** 2 + sigma ** 2)), math.log(1 + sigma ** 2 / mu ** 2)


def log_prior(params: dict, priors=None):
    """Sum of prior log-densities of the (positive) parameters; tensors may carry a leading chain axis."""
    priors = default_prior if priors is None else priors
    lp = 0.0
    for k, v in params.items():
        d = priors[k]
        lp = lp + type(d)(**{n: getattr(d, n).to(v.device, v.dtype) for n in d.arg_constraints}).log_prob(v)
    return lp

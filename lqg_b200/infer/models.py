"""Model glue -- counterpart of ``lqg/infer/models.py``: which constructor arguments are free parameters
(``get_model_params``, models.py:9-17) and the log-joint the inference loops differentiate.  The reference expresses the
models as numpyro programs; numpyro is not available here, so the log-joint is written out directly:

    log p(theta) + sum_trials log p(x_trial | theta)          (lqg_model / lifted_model, models.py:20-34,134)
"""
import inspect

import torch

from lqg_b200.infer.prior import default_prior, log_prior


def get_model_params(model_class):
    """Free parameters of a model class and their defaults (same exclusion list as the reference)."""
    sig = inspect.signature(model_class.__init__)
    skip = ["self", "dim", "dt", "T", "process_noise", "delay", "covar", "dtype", "device"]
    return {name: p.default for name, p in sig.parameters.items() if name not in skip}


def log_likelihood(theta: dict, x, model_type, process_noise=1.0, dt=1.0 / 60, dim=None, **fixed_params):
    """sum over trials of log p(x | theta).  ``theta`` values: tensors of shape () or (S,) (S parameter samples/chains)."""
    n, T1, d = x.shape
    kw = dict(process_noise=process_noise, dt=dt, T=T1 - 1, device=x.device, **fixed_params, **theta)
    if dim is not None:
        kw["dim"] = dim
    return model_type(**kw).log_likelihood(x).sum(-1)


def log_joint(theta: dict, x, model_type, priors=None, **kw):
    """Log-joint of ``lifted_model`` (models.py:134): default priors on the free parameters + likelihood."""
    return log_prior(theta, default_prior if priors is None else priors) + log_likelihood(theta, x, model_type, **kw)

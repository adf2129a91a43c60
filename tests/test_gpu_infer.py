"""GPU tests of the inference glue on top of the fused likelihood (counterpart of reference tests/infer_test.py:29-51)."""
import numpy as np
import pytest
import torch

from lqg_b200 import tracking
from lqg_b200.infer import get_model_params, infer, max_likelihood

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def test_get_model_params_matches_reference_lists():
    assert list(get_model_params(tracking.BoundedActor)) == ["action_variability", "sigma_target", "sigma_cursor", "action_cost"]
    assert len(get_model_params(tracking.SubjectiveActor)) == 6 and len(get_model_params(tracking.PointMassBoundedActor)) == 7
    assert len(get_model_params(tracking.OptimalActor)) == 3


def test_hmc_runs_and_returns_samples():
    """reference: infer(x, num_samples=10, num_warmup=10, model=BoundedActor); mcmc.get_samples() is not None"""
    T = 200
    x = tracking.BoundedActor(T=T, device=DEV).simulate(0, n=10)
    mcmc = infer(x, num_samples=10, num_warmup=10, model=tracking.BoundedActor, num_chains=8)
    s = mcmc.get_samples()
    assert set(s) == set(get_model_params(tracking.BoundedActor))
    assert all(v.shape == (80,) and torch.isfinite(v).all() and (v > 0).all() for v in s.values())


def test_max_likelihood_moves_towards_truth():
    T = 300
    true = dict(sigma_target=12.0, action_variability=0.5, action_cost=0.5, sigma_cursor=3.0)
    x = tracking.BoundedActor(T=T, device=DEV, **true).simulate(1, n=20)
    params, losses = max_likelihood(x, model=tracking.BoundedActor, steps=150, step_size=0.05, action_cost=0.5, sigma_cursor=3.0)
    assert losses[-1] < losses[0]
    assert abs(params["sigma_target"].item() - 12.0) < abs(6.0 - 12.0)      # moved from the default 6 towards 12

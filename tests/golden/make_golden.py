#!/usr/bin/env python
"""Generates tests/golden/*.npz: small golden input/output vectors of the hot path.

The reference (RothkopfLab/lqg) cannot be imported in this image (no jax / numpyro), so these vectors are
produced by the float64 oracle (oracle/lqg_np.py for values, oracle/lqg_torch.py autograd for parameter gradients), which
is itself pinned by tests/test_oracle_invariants.py.  They freeze the oracle against regressions and give the CUDA tests
fixed inputs.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import lqg_np as O  # noqa: E402
from oracle import lqg_torch as OT  # noqa: E402

CASES = {
    # name: (model, dim, T, N, params)
    "c1_bounded_T500": ("bounded", 1, 500, 20, dict(action_variability=0.5, sigma_target=6.0, sigma_cursor=6.0, action_cost=1.0)),
    "subjective2_T300": ("subjective", 2, 300, 6, dict(action_cost=0.7, action_variability=0.4, subj_noise=1.2, subj_vel_noise=0.6,
                                                       sigma_target=19.9, sigma_cursor=5.0)),
    "subjective1_T200": ("subjective", 1, 200, 5, dict(action_cost=1.0, action_variability=0.5, subj_noise=1.0, subj_vel_noise=0.5,
                                                       sigma_target=8.5, sigma_cursor=6.0)),
}


def main():
    for name, (model, dim, T, N, params) in CASES.items():
        np_builder = {"bounded": O.bounded_actor_mats, "subjective": O.subjective_actor_mats}[model]
        t_builder = {"bounded": OT.bounded_actor, "subjective": OT.subjective_actor}[model]
        mats = np_builder(dim=dim, **params)
        sa, sd = O.make_system(mats, T)
        X = O.simulate(sa, sd, N, np.random.default_rng(123)).astype(np.float32)
        ll = O.log_likelihood(sa, sd, X.astype(np.float64))
        L, _, H = O.lqr_backward(sa)
        K = O.kf_forward(sa, sa["V"][0] @ sa["V"][0].T)
        names = sorted(params)
        th = [torch.tensor(params[k], dtype=torch.float64, requires_grad=True) for k in names]
        a, d = t_builder(dim=dim, **dict(zip(names, th)))
        llt = OT.log_likelihood(a, d, torch.tensor(X, dtype=torch.float64))
        g = torch.autograd.grad(llt.sum(), th)
        assert np.allclose(llt.detach().numpy(), ll, rtol=1e-10)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), model=model, dim=dim, T=T, N=N, param_names=np.array(names),
                            param_values=np.array([params[k] for k in names]), X=X, ll=ll,
                            grad=np.array([gi.item() for gi in g]), L_first=L[0], L_last=L[-1], K_first=K[0], K_last=K[-1],
                            H_first=H[0])
        print(name, "ll sum", ll.sum(), "grad", [round(gi.item(), 4) for gi in g])


if __name__ == "__main__":
    main()

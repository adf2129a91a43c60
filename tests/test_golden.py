"""Golden-vector tests (tests/golden/*.npz, produced by tests/golden/make_golden.py from the float64 oracle).

not-gpu: the oracle still reproduces them; the kernels' step functions (host emulation, through the C ABI structs and
the torch model builders) match them.  gpu: the CUDA library through the public Python API matches them."""
import glob
import os

import numpy as np
import pytest
import torch

from lqg_b200 import abi, tracking
from oracle import lqg_np as O
from tests import helpers as H

FIXTURES = sorted(p for p in glob.glob(os.path.join(H.ROOT, "tests", "golden", "*.npz")) if not os.path.basename(p).startswith("ref_"))   # ref_*: tests/test_reference_golden.py
CLS = {"bounded": tracking.BoundedActor, "subjective": tracking.SubjectiveActor}
NPB = {"bounded": O.bounded_actor_mats, "subjective": O.subjective_actor_mats}
GRAD_ATOL_FRAC = 1e-5     # absolute slack as a fraction of max |grad| (tiny components near an optimum)


def _load(path):
    z = np.load(path)
    params = dict(zip([str(k) for k in z["param_names"]], [float(v) for v in z["param_values"]]))
    return z, str(z["model"]), int(z["dim"]), int(z["T"]), int(z["N"]), params


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_oracle_reproduces_golden(path):
    z, model, dim, T, N, params = _load(path)
    sa, sd = O.make_system(NPB[model](dim=dim, **params), T)
    ll = O.log_likelihood(sa, sd, z["X"].astype(np.float64))
    assert np.allclose(ll, z["ll"], rtol=1e-12)
    L, _, Hh = O.lqr_backward(sa)
    K = O.kf_forward(sa, sa["V"][0] @ sa["V"][0].T)
    assert np.allclose(L[0], z["L_first"], rtol=1e-12) and np.allclose(K[-1], z["K_last"], rtol=1e-12)
    assert np.allclose(Hh[0], z["H_first"], rtol=1e-12)


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_step_functions_reproduce_golden(path):
    """Host emulation of the kernels' step functions + torch model builders + chain rule to the parameters."""
    z, model, dim, T, N, params = _load(path)
    lib = abi.Library(H.EMUL_PATH)
    names = sorted(params)
    th = [torch.tensor(params[k], dtype=torch.float64, requires_grad=True) for k in names]
    m = CLS[model](dim=dim, T=T, device="cpu", dtype=torch.float64, **dict(zip(names, th)))
    x, b, u, y = m.xdim, m.bdim, m.udim, m.ydim
    dims = abi.LqgkDims(1, N, T, x, b, u, y, x)
    act = {k: getattr(m.actor, k)[0].contiguous()[None] for k in abi.ACTOR_KEYS}
    dyn = {k: getattr(m.dynamics, k)[0].contiguous()[None] for k in abi.DYN_KEYS}
    x_tm = lib.pack_obs(torch.tensor(z["X"]))
    ll, ga, gd, _ = lib.loglik_vjp(dims, {k: v.detach() for k, v in act.items()}, {k: v.detach() for k, v in dyn.items()}, x_tm)
    assert np.allclose(ll[0].numpy(), z["ll"], rtol=1e-4)
    outs = [act[k] for k in abi.ACTOR_KEYS] + [dyn[k] for k in abi.DYN_KEYS]
    cots = [ga[k] for k in abi.ACTOR_KEYS] + [gd[k] for k in abi.DYN_KEYS]
    keep = [(o, c) for o, c in zip(outs, cots) if o.requires_grad]
    g = torch.autograd.grad([o for o, _ in keep], th, grad_outputs=[c for _, c in keep], allow_unused=True)
    g = np.array([0.0 if gi is None else gi.item() for gi in g])
    assert np.allclose(g, z["grad"], rtol=1e-3, atol=GRAD_ATOL_FRAC * np.abs(z["grad"]).max()), (g, z["grad"])
    L, _, Hh = lib.lqr_backward(dims, {k: v.detach() for k, v in act.items()})
    K = lib.kf_forward(dims, {k: v.detach() for k, v in act.items()})
    assert np.allclose(L[0, 0].numpy(), z["L_first"], rtol=1e-9, atol=1e-12) and np.allclose(L[0, -1].numpy(), z["L_last"], rtol=1e-9, atol=1e-12)
    assert np.allclose(K[0, 0].numpy(), z["K_first"], rtol=1e-9, atol=1e-12) and np.allclose(K[0, -1].numpy(), z["K_last"], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_cuda_public_api_reproduces_golden(path, dtype):
    z, model, dim, T, N, params = _load(path)
    dev = torch.device("cuda:0")
    names = sorted(params)
    th = [torch.tensor(params[k], dtype=dtype, device=dev, requires_grad=True) for k in names]
    m = CLS[model](dim=dim, T=T, **dict(zip(names, th)))
    ll = m.log_likelihood(torch.tensor(z["X"], device=dev))
    ll.sum().backward()
    assert np.allclose(ll.detach().cpu().numpy(), z["ll"], rtol=1e-4)
    g = np.array([t.grad.item() for t in th])
    assert np.allclose(g, z["grad"], rtol=1e-3, atol=GRAD_ATOL_FRAC * np.abs(z["grad"]).max()), (g, z["grad"])

#!/usr/bin/env python
"""Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES (read from /root/reference, unmodified).

jax / numpyro cannot be installed in this image, so the reference's modules are imported over `tests/golden/refshim`, a
float64 NumPy stand-in for the few jax / numpyro names they use (see refshim/README.md).  What runs is the reference's
lqg/spec.py, lqg/utils.py, lqg/control/lqr.py, lqg/belief/kf.py, lqg/system.py and lqg/tracking/*.py, byte for byte;
`lqg/__init__.py` is bypassed (it asks importlib.metadata for an installed distribution) by registering an empty `lqg`
package object whose __path__ points at /root/reference/lqg.

Per case the fixture holds
  model / kwargs          reference class name and constructor arguments (differentiated parameters separately)
  actor_*, dyn_*          the 12 base matrices the reference constructor produced (time slice 0)
  jac_actor_*, jac_dyn_*  d(base matrix)/d(parameter) through the reference constructor (central differences, [P,...])
  X                       trajectories drawn by the reference's System.simulate (NumPy PRNG stream), rounded to float32
  ll                      System.log_likelihood(X) per trial (float64)
  grad, grad_err          d sum(ll) / d parameter by Richardson-extrapolated central differences THROUGH THE REFERENCE
                          (what jax.grad would return up to the quoted error estimate)
  L, l_absmax, H, K       lqr.backward / kf.forward outputs
  mu_*, Sigma_*           conditional_moments of trial 0 at a few steps

Run here only (needs /root/reference):  python tests/golden/make_reference_golden.py [case ...]
"""
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    sys.path.insert(0, os.path.join(HERE, "refshim"))
    pkg = types.ModuleType("lqg")
    pkg.__path__ = [os.path.join(REF, "lqg")]
    sys.modules["lqg"] = pkg
    import lqg.system  # noqa: F401
    import lqg.tracking
    import lqg.tracking.delay
    from lqg.belief import kf
    from lqg.control import lqr
    from jax import random
    return lqg.tracking, lqg.tracking.delay, lqr, kf, random


# name: (class, fixed kwargs, differentiated parameters, delay, observed dims, trials, PRNG seed)
CASES = {
    # tests/infer_test.py:18-26 (BoundedActor(T=500), PRNGKey(123), n=20) = BASELINE config c1
    "ref_c1_bounded_T500": ("BoundedActor", dict(T=500), dict(action_variability=0.5, sigma_target=6.0, sigma_cursor=6.0, action_cost=1.0),
                            0, None, 20, 123),
    # tests/infer_test.py:10-16 (SubjectiveActor(T=500), PRNGKey(113), n=20)
    "ref_subjective_T500": ("SubjectiveActor", dict(T=500), dict(action_cost=1.0, action_variability=0.5, subj_noise=1.0, subj_vel_noise=0.5,
                                                                  sigma_target=6.0, sigma_cursor=6.0), 0, None, 20, 113),
    "ref_subjective2_T300": ("SubjectiveActor", dict(dim=2, T=300), dict(action_cost=0.7, action_variability=0.4, subj_noise=1.2,
                                                                          subj_vel_noise=0.6, sigma_target=19.9, sigma_cursor=5.0), 0, None, 6, 5),
    # BASELINE config c2/c3 shape: SubjectiveActor 2-D, T=1200, 20 trials of one blob-width condition
    "ref_c2_subjective2_T1200": ("SubjectiveActor", dict(dim=2, T=1200), dict(action_cost=1.0, action_variability=0.5, subj_noise=1.0,
                                                                               subj_vel_noise=0.5, sigma_target=19.9, sigma_cursor=6.0), 0, None, 20, 7),
    "ref_bounded2_T200": ("BoundedActor", dict(dim=2, T=200), dict(action_variability=0.6, sigma_target=9.0, sigma_cursor=4.0, action_cost=0.5),
                          0, None, 5, 21),
    "ref_optimal_T200": ("OptimalActor", dict(T=200), dict(action_variability=0.5, sigma_target=6.0, sigma_cursor=6.0), 0, None, 5, 3),
    "ref_relobs_T200": ("RelativeObservationBoundedActor", dict(T=200), dict(action_variability=0.5, sigma=6.0, action_cost=1.0), 0, None, 5, 4),
    # main.py:51 use of the point-mass model: only target and cursor position are observed
    "ref_pointmass_T200": ("PointMassBoundedActor", dict(T=200), dict(action_variability=1e-3, sigma_target=6.0, sigma_cursor=6.0,
                                                                        action_cost=0.01), 0, 2, 5, 9),
    # motor-delay family (tracking/delay.py:36-41) on the bounded actor, delay of 2 steps, positions observed
    "ref_delay2_bounded_T150": ("BoundedActor", dict(T=150), dict(action_variability=0.5, sigma_target=6.0, sigma_cursor=6.0, action_cost=1.0),
                                2, 2, 5, 13),
    # BASELINE config c4: TemporalDelayModel(PointMassBoundedActor(T=600), delay=2) -- 12-dim state, joint dim 24
    "ref_c4_pmdelay2_T600": ("PointMassBoundedActor", dict(T=600), dict(action_variability=1e-3, sigma_target=6.0, sigma_cursor=6.0,
                                                                          action_cost=0.01), 2, 2, 4, 13),
    # BASELINE config c2r: the REAL tracking data of the reference repository (data/data.mat, Bonnen et al. 2015) through
    # the reference's own loader lqg/io.py:45-98 (delay=12, clip=120), narrowest-blob condition: 20 trials x 1068 samples x (target,
    # cursor); BoundedActor dim=1 at sigma_target 8.5 (the notebook's posterior mean for that condition)
    "ref_c2r_bounded_realdata_T1067": ("BoundedActor", dict(T=1067), dict(action_variability=0.5, sigma_target=8.5, sigma_cursor=6.0,
                                                                          action_cost=1.0), 0, None, 20, "realdata:0"),
}

ACTOR_KEYS = ("A", "B", "F", "V", "W", "Q", "R")
DYN_KEYS = ("A", "B", "F", "V", "W")


def T_of(fixed):
    return int(fixed["T"])


def main(argv):
    tracking, delay_mod, lqr, kf, random = import_reference()

    def build(cls, fixed, params, delay):
        m = getattr(tracking, cls)(**fixed, **params)
        return delay_mod.TemporalDelayModel(m, delay) if delay else m

    def mats(m):
        out = {"actor_" + k: np.asarray(getattr(m.actor, k)[0], dtype=np.float64) for k in ACTOR_KEYS}
        out.update({"dyn_" + k: np.asarray(getattr(m.dynamics, k)[0], dtype=np.float64) for k in DYN_KEYS})
        return out

    for name, (cls, fixed, params, delay, obs, N, seed) in CASES.items():
        if argv and name not in argv:
            continue
        t0 = time.time()
        m = build(cls, fixed, params, delay)
        if isinstance(seed, str) and seed.startswith("realdata:"):
            import importlib.util
            spec = importlib.util.spec_from_file_location("lqg_io", os.path.join(REF, "lqg", "io.py"))
            io = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(io)
            data, sigmas = io.load_tracking_data(delay=12, clip=120, data_path=os.path.join(REF, "data"))
            X = np.asarray(data[int(seed.split(":")[1])][:N], dtype=np.float64)
            assert X.shape == (N, T_of(fixed) + 1, 2), X.shape
        else:
            X = np.asarray(m.simulate(random.PRNGKey(seed), n=N))
        if obs is not None:
            X = X[..., :obs]
        X = X.astype(np.float32)
        X64 = np.asarray(X, dtype=np.float64)

        def sum_ll(p):
            return float(np.sum(build(cls, fixed, p, delay).log_likelihood(X64)))

        ll = np.asarray(m.log_likelihood(X64), dtype=np.float64)
        gains = lqr.backward(m.actor)
        K = kf.forward(m.actor, Sigma0=m.actor.V[0] @ m.actor.V[0].T)
        mu, Sig = m.conditional_moments(X64[0])
        T = int(fixed["T"])
        steps = np.array([0, 1, T // 2, T - 1])

        names = sorted(params)
        grad, gerr = np.zeros(len(names)), np.zeros(len(names))
        base = mats(m)
        jac = {k: np.zeros((len(names),) + v.shape) for k, v in base.items()}
        for i, k in enumerate(names):
            h = 1e-3 * abs(params[k])
            d = []
            for hh in (h, h / 2):
                pp, pm = dict(params), dict(params)
                pp[k] += hh
                pm[k] -= hh
                d.append((sum_ll(pp) - sum_ll(pm)) / (2 * hh))
            grad[i] = (4 * d[1] - d[0]) / 3            # Richardson: O(h^4)
            gerr[i] = abs(d[1] - d[0]) / 3
            hj = 1e-4 * abs(params[k])
            pp, pm = dict(params), dict(params)
            pp[k] += hj
            pm[k] -= hj
            mp, mm = mats(build(cls, fixed, pp, delay)), mats(build(cls, fixed, pm, delay))
            for kk in jac:
                jac[kk][i] = (mp[kk] - mm[kk]) / (2 * hj)

        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), model=cls, delay=delay, obs_dim=X.shape[-1], T=T, N=N,
            fixed_names=np.array(sorted(fixed)), fixed_values=np.array([fixed[k] for k in sorted(fixed)]),
            param_names=np.array(names), param_values=np.array([params[k] for k in names]),
            X=X, ll=ll, grad=grad, grad_err=gerr, L=np.asarray(gains.L), l_absmax=float(np.abs(gains.l).max()), H=np.asarray(gains.H),
            K=np.asarray(K), steps=steps, mu_steps=np.asarray(mu)[steps], Sigma_steps=np.asarray(Sig)[steps],
            **base, **{"jac_" + k: v for k, v in jac.items()})
        print(f"{name}: sum ll {ll.sum():.6f}  grad {np.round(grad, 4)}  rel FD err {np.max(gerr / (np.abs(grad).max())):.1e}  "
              f"({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])

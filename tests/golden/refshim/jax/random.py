"""`jax.random` stand-in: keys are NumPy SeedSequences (streams differ from threefry; data are stored in the fixtures)."""
import numpy as _np

from .numpy import Array


class _Key:
    def __init__(self, ss):
        self.ss = ss

    def __len__(self):
        return 2


def PRNGKey(seed):
    return _Key(_np.random.SeedSequence(int(seed)))


key = PRNGKey


class _KeyList(list):
    pass


def split(k, num=2):
    return _KeyList(_Key(s) for s in k.ss.spawn(num))


def normal(k, shape=(), dtype=None):
    return _np.random.Generator(_np.random.PCG64(k.ss)).standard_normal(shape).view(Array)

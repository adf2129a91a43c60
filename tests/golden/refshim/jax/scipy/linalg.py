"""`jax.scipy.linalg` stand-in (block_diag, expm, cholesky) over SciPy, float64."""
import numpy as _np
import scipy.linalg as _sl

from ..numpy import Array


def block_diag(*arrs):
    return _np.asarray(_sl.block_diag(*[_np.asarray(a, dtype=_np.float64) for a in arrs]), dtype=_np.float64).view(Array)


def expm(a):
    return _np.asarray(_sl.expm(_np.asarray(a, dtype=_np.float64))).view(Array)


def cholesky(a, lower=False):
    return _np.asarray(_sl.cholesky(_np.asarray(a, dtype=_np.float64), lower=lower)).view(Array)

"""`numpyro.distributions` stand-in: Distribution base, MultivariateNormal with log_prob/to_event/shape (Cholesky density)."""
import numpy as _np


class Distribution:
    def __init__(self, batch_shape=(), event_shape=()):
        self.batch_shape, self.event_shape = tuple(batch_shape), tuple(event_shape)

    def shape(self, sample_shape=()):
        return tuple(sample_shape) + self.batch_shape + self.event_shape

    def to_event(self, n=None):
        return Independent(self, len(self.batch_shape) if n is None else n)


class MultivariateNormal(Distribution):
    def __init__(self, loc, covariance_matrix):
        self.loc = _np.asarray(loc, dtype=_np.float64)
        self.covariance_matrix = _np.asarray(covariance_matrix, dtype=_np.float64)
        bs = _np.broadcast_shapes(self.loc.shape[:-1], self.covariance_matrix.shape[:-2])
        super().__init__(batch_shape=bs, event_shape=self.loc.shape[-1:])
        self.scale_tril = _np.linalg.cholesky(self.covariance_matrix)

    def log_prob(self, value):
        d = self.loc.shape[-1]
        diff = _np.asarray(value, dtype=_np.float64) - self.loc
        Lc = _np.broadcast_to(self.scale_tril, diff.shape[:-1] + (d, d))
        z = _np.linalg.solve(Lc, diff[..., None])[..., 0]       # L z = diff (general solve of a triangular system)
        half_logdet = _np.log(_np.diagonal(Lc, axis1=-2, axis2=-1)).sum(-1)
        return -0.5 * (z * z).sum(-1) - half_logdet - 0.5 * d * _np.log(2 * _np.pi)


class Independent(Distribution):
    def __init__(self, base, n):
        self.base_dist, self.n = base, n
        bs = base.batch_shape
        super().__init__(batch_shape=bs[:len(bs) - n], event_shape=bs[len(bs) - n:] + base.event_shape)

    def log_prob(self, value):
        lp = self.base_dist.log_prob(value)
        return lp.sum(axis=tuple(range(-self.n, 0))) if self.n else lp

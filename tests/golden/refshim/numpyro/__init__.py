"""`numpyro` stand-in: only `numpyro.distributions.{Distribution, MultivariateNormal}` (see ../README.md)."""
from . import distributions  # noqa: F401

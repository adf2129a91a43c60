"""JAX binding of liblqgk.so through the XLA FFI (custom call + ``jax.custom_vjp``).

UNTESTED IN THIS ENVIRONMENT (no jax / jaxlib in the image or wheelhouse, no network).  Importing this module raises
``ImportError`` unless jax is installed *and* ``lqg_b200/xla/liblqgk_xla.so`` has been built from ``lqgk_xla_ffi.cc``
(recipe in that file).  It is what makes ``lqg.system.System.log_likelihood`` of the reference a drop-in:

    from lqg_b200.xla.jax_binding import log_likelihood          # (actor: LQGSpec, dynamics: LQGSpec, x) -> ll[n]
    class System: ...
        def log_likelihood(self, x, Sigma0=None):                # lqg/system.py:246-248
            return log_likelihood(self.actor, self.dynamics, x)

Batching over parameters: pass specs with a leading sample axis ``[S, T, r, c]`` (the kernels' own sample axis).  ``jax.vmap``
of this function is only supported as a sequential loop over the mapped axis (``vmap_method="sequential"``): the C++ handlers
read (S, T, N, d) from fixed operand positions, so a batch axis prepended by vmap cannot be folded into S there.
Limits (not checkable under tracing, so stated here): time-invariant specs only -- ``_base`` takes the t = 0 slice of every
stacked array, like the reference models build them (lqg/utils.py:6-35) --, and the default ``Sigma0`` (= V_a[0] V_a[0]^T,
lqg/system.py:158-161); anything else must go through the C ABI directly (include/lqgk.h).
"""
import ctypes
import os

try:
    import jax
    import jax.numpy as jnp
    import numpy as np
except ImportError as e:  # pragma: no cover - jax is not available in this image
    raise ImportError("lqg_b200.xla.jax_binding needs jax/jaxlib (not installable in this image); "
                      "use the torch-facing API lqg_b200.system / lqg_b200.tracking instead") from e

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = ctypes.CDLL(os.path.join(_HERE, "liblqgk_xla.so"))
_ABI = ctypes.CDLL(os.path.join(_HERE, "..", "csrc", "liblqgk.so"))
jax.ffi.register_ffi_target("lqg_loglik_fwd", jax.ffi.pycapsule(_LIB.LqgLoglikFwd), platform="CUDA")
jax.ffi.register_ffi_target("lqg_loglik_vjp", jax.ffi.pycapsule(_LIB.LqgLoglikVjp), platform="CUDA")

ACT = ("A", "B", "F", "V", "W", "Q", "R")
DYN = ("A", "B", "F", "V", "W")


def _base(M):
    """[T, r, c] time-stacked (time-invariant) -> [1, r, c]; [S, T, r, c] -> [S, r, c]."""
    return M[..., 0, :, :].reshape((-1,) + M.shape[-2:]).astype(jnp.float32)


def _workspace_bytes(S, N, T, x, b, u, y, d, mode):
    class Dims(ctypes.Structure):
        _fields_ = [(k, ctypes.c_int32) for k in ("S", "N", "T", "x", "b", "u", "y", "d")] + [("x_sample_stride", ctypes.c_int64)]
    _ABI.lqgk_workspace_bytes.restype = ctypes.c_size_t
    return int(_ABI.lqgk_workspace_bytes(ctypes.byref(Dims(S, N, T, x, b, u, y, d, 0)), mode, 8192))


def _mats(actor, dynamics):
    return [_base(getattr(actor, k)) for k in ACT] + [_base(getattr(dynamics, k)) for k in DYN]


@jax.custom_vjp
def _loglik(mats, x_tm):
    return _fwd_call(mats, x_tm)


def _shapes(mats, x_tm):
    S = max(m.shape[0] for m in mats)
    T1, N, d = x_tm.shape
    x, b, u, y = mats[7].shape[-1], mats[0].shape[-1], mats[1].shape[-1], mats[2].shape[-2]
    return S, N, T1 - 1, x, b, u, y, d


def _fwd_call(mats, x_tm):
    S, N, T, x, b, u, y, d = _shapes(mats, x_tm)
    out = (jax.ShapeDtypeStruct((S, N), jnp.float32),
           jax.ShapeDtypeStruct((_workspace_bytes(S, N, T, x, b, u, y, d, 1),), jnp.uint8))
    ll, _ = jax.ffi.ffi_call("lqg_loglik_fwd", out, vmap_method="sequential")(*mats, x_tm)
    return ll


def _vjp_fwd(mats, x_tm):
    return _fwd_call(mats, x_tm), (mats, x_tm)


def _vjp_bwd(res, ll_bar):
    mats, x_tm = res
    S, N, T, x, b, u, y, d = _shapes(mats, x_tm)
    outs = ([jax.ShapeDtypeStruct((S, N), jnp.float32)]
            + [jax.ShapeDtypeStruct((S,) + m.shape[1:], jnp.float32) for m in mats]
            + [jax.ShapeDtypeStruct((_workspace_bytes(S, N, T, x, b, u, y, d, 2),), jnp.uint8)])
    res = jax.ffi.ffi_call("lqg_loglik_vjp", tuple(outs), vmap_method="sequential")(*mats, x_tm, ll_bar.astype(jnp.float32))
    grads = [g if m.shape[0] > 1 else g.sum(0, keepdims=True) for g, m in zip(res[1:13], mats)]
    return grads, None


_loglik.defvjp(_vjp_fwd, _vjp_bwd)


def log_likelihood(actor, dynamics, x):
    """Drop-in for ``System.log_likelihood`` (lqg/system.py:246-248): x[n, T+1, d] -> ll[n] (or [S, n])."""
    x_tm = jnp.transpose(x.astype(jnp.float32), (1, 0, 2))
    ll = _loglik(_mats(actor, dynamics), x_tm)
    return ll[0] if actor.A.ndim == 3 else ll

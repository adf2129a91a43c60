"""CPU tests of the host-side logic that needs no GPU: model containers, `.to()`, spec bookkeeping."""
import torch

from lqg_b200 import runtime, tracking
from lqg_b200.utils import is_known_zero, time_stack_spec


def test_to_converts_the_cached_axis_system():
    """BoundedActor(dim=2).to(float64) left a float32 `_axis_system` behind."""
    for cls in (tracking.BoundedActor, tracking.SubjectiveActor, tracking.RelativeObservationBoundedActor):
        m = cls(dim=2, T=20, device="cpu")
        m64 = m.to(torch.float64)
        assert m64.dtype == torch.float64
        assert m64._axis_system.dtype == torch.float64
        assert m64._axis_system.actor.V.dtype == torch.float64
        assert m._axis_system.dtype == torch.float32          # the original is untouched


def test_default_terminal_cost_and_zero_cross_cost_are_recognised_without_reading_values():
    I = torch.eye(2)
    sp = time_stack_spec(I, I[:, :1], I, I, I, I, I[:1, :1], T=5)
    assert runtime._qf_is_default(sp) and is_known_zero(sp.P)
    assert not runtime._qf_is_default(sp._replace(Qf=2 * I))
    S = 3
    spb = time_stack_spec(I.expand(S, 2, 2), I[:, :1], I, I, I, I, I[:1, :1] * torch.arange(1.0, 4.0)[:, None, None], T=5)
    assert runtime._qf_is_default(spb) and is_known_zero(spb.P)
    assert not is_known_zero(spb.P.clone())


def test_dimension_tuple_lists_are_parsed_from_the_header():
    """lqg_b200.dims reads csrc/lqgk_dims.h (one source of truth): the tuples of the main path and the ones compiled for the
    all-FP64 per-trial likelihood only."""
    from lqg_b200 import dims
    assert (2, 3, 1, 2, 2) in dims.SUPPORTED_DIMS and (12, 12, 1, 3, 2) in dims.SUPPORTED_DIMS and len(dims.SUPPORTED_DIMS) == 10
    assert dims.FP64_ONLY_DIMS == [(4, 4, 1, 3, 4)] and not set(dims.FP64_ONLY_DIMS) & set(dims.SUPPORTED_DIMS)

// lqgk_dims.h -- the compiled (x, b, u, y, d) instantiations: dynamics state, belief state, control,
// observation, observed-by-experimenter dims.  One line per reference model family (lqg/tracking/*.py).
#pragma once
// clang-format off
#define LQGK_FOR_EACH_DIMS(M)                                                                       \
  M(2, 2, 1, 2, 2)   /* BoundedActor / OptimalActor dim=1 (basic.py:43-87)            [config c1] */ \
  M(2, 2, 1, 1, 2)   /* RelativeObservationBoundedActor dim=1 (basic.py:90-124)                   */ \
  M(2, 3, 1, 2, 2)   /* SubjectiveActor dim=1 (subjective.py:15-47)                               */ \
  M(4, 4, 2, 4, 4)   /* BoundedActor dim=2                                                        */ \
  M(4, 4, 2, 2, 4)   /* RelativeObservationBoundedActor dim=2                                     */ \
  M(4, 6, 2, 4, 4)   /* SubjectiveActor dim=2                               [configs c2, c3, c5] */ \
  M(4, 4, 1, 3, 2)   /* PointMassBoundedActor, observed target+cursor (point_mass.py, main.py:51) */ \
  M(5, 5, 1, 2, 2)   /* HandMotionModelTrackingTask (notebooks/HandModel.ipynb): target + pos/vel/force/act. */ \
  M(6, 6, 1, 2, 2)   /* TemporalDelayModel(BoundedActor dim=1, delay=2) (delay.py:9-41): motor-delay family */ \
  M(12, 12, 1, 3, 2) /* TemporalDelayModel(PointMassBoundedActor, delay=2): 12-dim state, joint dim 24 [config c4]; large-system path (lqgk_big.cuh) */
// clang-format on

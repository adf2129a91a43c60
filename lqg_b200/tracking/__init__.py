from lqg_b200.tracking.basic import BoundedActor, OptimalActor, RelativeObservationBoundedActor
from lqg_b200.tracking.point_mass import PointMassBoundedActor
from lqg_b200.tracking.subjective import SubjectiveActor
from lqg_b200.tracking.delay import DelayedSubjectiveActor, TemporalDelayModel
from lqg_b200.tracking.hand import HandMotionModelTrackingTask

__all__ = ["BoundedActor", "OptimalActor", "RelativeObservationBoundedActor", "SubjectiveActor",
           "PointMassBoundedActor", "TemporalDelayModel", "DelayedSubjectiveActor", "HandMotionModelTrackingTask"]

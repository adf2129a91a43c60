from lqg_b200.tracking.basic import BoundedActor, OptimalActor, RelativeObservationBoundedActor
from lqg_b200.tracking.point_mass import PointMassBoundedActor
from lqg_b200.tracking.subjective import SubjectiveActor

__all__ = ["BoundedActor", "OptimalActor", "RelativeObservationBoundedActor", "SubjectiveActor",
           "PointMassBoundedActor"]

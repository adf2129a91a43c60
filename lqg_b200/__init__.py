"""lqg_b200 -- B200-native (sm_100a) implementation of the LQG inverse-optimal-control likelihood path of
RothkopfLab/lqg, behind the reference's Python API.  See DESIGN.md."""
__version__ = "0.1.0"

from lqg_b200.spec import LQGSpec  # noqa: E402,F401
from lqg_b200.system import LQG, Actor, Dynamics, System  # noqa: E402,F401

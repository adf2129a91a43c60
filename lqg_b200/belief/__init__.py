from lqg_b200.belief import kf  # noqa: F401

from lqg_b200.control import lqr  # noqa: F401

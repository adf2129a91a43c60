"""Compiled (x, b, u, y, d) dimension tuples -- parsed from csrc/lqgk_dims.h so there is one source of truth."""
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))


def _parse():
    txt = open(os.path.join(_HERE, "csrc", "lqgk_dims.h")).read()
    return [tuple(int(v) for v in m) for m in re.findall(r"M\((\d+), (\d+), (\d+), (\d+), (\d+)\)", txt)]


SUPPORTED_DIMS = _parse()

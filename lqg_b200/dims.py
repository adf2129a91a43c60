"""Compiled (x, b, u, y, d) dimension tuples -- parsed from csrc/lqgk_dims.h so there is one source of truth."""
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))


def _parse(macro):
    txt = open(os.path.join(_HERE, "csrc", "lqgk_dims.h")).read()
    body = txt.split("#define " + macro + "(M)", 1)[1].split("#define", 1)[0].split("// clang-format on", 1)[0]
    return [tuple(int(v) for v in m) for m in re.findall(r"M\((\d+), (\d+), (\d+), (\d+), (\d+)\)", body)]


SUPPORTED_DIMS = _parse("LQGK_FOR_EACH_DIMS")
FP64_ONLY_DIMS = _parse("LQGK_FOR_EACH_FP64_ONLY_DIMS")   # all-FP64 per-trial likelihood only (System.log_likelihood_fp64)

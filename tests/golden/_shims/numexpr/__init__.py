"""numexpr stand-in: evaluate the expression with numpy in the caller's frame (golden generation only)."""
import sys
import numpy as np


def evaluate(expr, local_dict=None, global_dict=None):
    frame = sys._getframe(1)
    ns = {k: getattr(np, k) for k in ("exp", "sin", "cos", "sqrt", "log", "abs", "arctan2", "where")}
    ns.update(frame.f_globals if global_dict is None else global_dict)
    ns.update(frame.f_locals if local_dict is None else local_dict)
    ns.update({k: getattr(np, k) for k in ("exp", "sin", "cos", "sqrt", "log")})
    return eval(expr, {"__builtins__": {}}, ns)

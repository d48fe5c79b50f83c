"""Alias of rvgp_b200.smoothing (drop-in module path of the reference's RVGP/smoothing.py)."""
from rvgp_b200.smoothing import *  # noqa: F401,F403
from rvgp_b200 import smoothing as _m


def __getattr__(name):
    return getattr(_m, name)

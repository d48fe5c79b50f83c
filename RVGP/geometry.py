"""Alias of rvgp_b200.geometry (drop-in module path of the reference's RVGP/geometry.py)."""
from rvgp_b200.geometry import *  # noqa: F401,F403
from rvgp_b200 import geometry as _m


def __getattr__(name):
    return getattr(_m, name)

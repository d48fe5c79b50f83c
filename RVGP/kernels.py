"""Alias of rvgp_b200.kernels (drop-in module path of the reference's RVGP/kernels.py)."""
from rvgp_b200.kernels import *  # noqa: F401,F403
from rvgp_b200 import kernels as _m


def __getattr__(name):
    return getattr(_m, name)

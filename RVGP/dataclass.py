"""Alias of rvgp_b200.dataclass (drop-in module path of the reference's RVGP/dataclass.py)."""
from rvgp_b200.dataclass import *  # noqa: F401,F403
from rvgp_b200 import dataclass as _m


def __getattr__(name):
    return getattr(_m, name)

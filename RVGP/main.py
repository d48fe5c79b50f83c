"""Alias of rvgp_b200.main (drop-in module path of the reference's RVGP/main.py)."""
from rvgp_b200.main import *  # noqa: F401,F403
from rvgp_b200 import main as _m


def __getattr__(name):
    return getattr(_m, name)

"""Drop-in alias of the reference's RVGP/utils.py."""
from rvgp_b200.utils import load_mesh  # noqa: F401

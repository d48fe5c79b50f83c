"""rvgp_b200: B200-native (sm_100a) implementation of RVGP's geometry-and-spectral hot path.

Host-side mirror of the reference's Python surface over the C-ABI library librvgp_b200.so.
"""
__version__ = "0.1.0"

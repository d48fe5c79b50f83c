"""CPU-only: the C-ABI library loads and exports every symbol include/rvgp_b200.h declares."""
import ctypes

from rvgp_b200 import _cabi


def test_library_exports_all_declared_symbols():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = _cabi.declared_symbols()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.rvgp_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        _cabi.Handle(0)

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


def test_hot_spmm_kernel_does_not_spill():
    """The d=2 / 64-column fused SpMM instantiation is register-capped (6 CTAs/SM); extra kernel parameters once pushed
    it into local-memory spills and cost 25 % -- keep that from regressing silently (ptxas -v log of the build)."""
    import os
    import re
    import __graft_entry__ as g
    g.build()
    log = os.path.join(os.path.dirname(_cabi.LIB_PATH), "..", "build", "spmm.o.log")
    txt = open(log).read()
    i = txt.find("Function properties for _ZN4rvgp18bsr_spmm_v2_kernelILi2ELi32ELi1ELi4ELb0ELb0ELb0ELb0E")
    assert i >= 0, "hot instantiation not found in the ptxas log"
    chunk = txt[i:i + 600]
    assert "0 bytes spill stores, 0 bytes spill loads" in chunk, chunk
    regs = int(re.search(r"Used (\d+) registers", chunk).group(1))
    assert regs <= 40


def test_hot_mma_spmm_kernel_does_not_spill():
    """The shipped FP64-MMA SpMM instantiation (4 chunks, ring depth 2, 256 threads x 2 CTAs/SM, compact fragments) must
    stay spill-free: on B200 the spilling variants of the same kernel ran 1.3-1.9 ms instead of 0.83 ms."""
    import os
    import re
    import __graft_entry__ as g
    g.build()
    log = os.path.join(os.path.dirname(_cabi.LIB_PATH), "..", "build", "spmm_mma.o.log")
    txt = open(log).read()
    # AMODE 1 = compact rotations (connection Laplacian), AMODE 2 = pattern mode (scalar Laplacian as L (x) I_2)
    for amode in (1, 2):
        i = txt.find("Function properties for _ZN4rvgp26bsr_spmm_mma_native_kernelILi4ELi2ELi256ELi2ELi%dE" % amode)
        assert i >= 0, "hot instantiation (AMODE %d) not found in the ptxas log" % amode
        chunk = txt[i:i + 600]
        assert "0 bytes spill stores, 0 bytes spill loads" in chunk, chunk
        assert int(re.search(r"Used (\d+) registers", chunk).group(1)) <= 128

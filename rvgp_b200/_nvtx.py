"""NVTX ranges per pipeline stage: `nsys` / `ncu --nvtx` timelines show the stages of
create_data_object / fit / transform and the phases of the block eigensolver by name.  torch.cuda.nvtx is the binding (the
ranges cost ~100 ns each when no profiler is attached); RVGP_NVTX=0 turns them into no-ops."""
import contextlib
import os

_ON = os.environ.get("RVGP_NVTX", "1") != "0"
_nvtx = None


def _lib():
    global _nvtx, _ON
    if _nvtx is None and _ON:
        try:
            import torch.cuda.nvtx as nv
            nv.range_push("rvgp_b200")
            nv.range_pop()
            _nvtx = nv
        except Exception:
            _ON = False
    return _nvtx if _ON else None


@contextlib.contextmanager
def stage(name):
    nv = _lib()
    if nv is None:
        yield
        return
    nv.range_push("rvgp:" + name)
    try:
        yield
    finally:
        nv.range_pop()


def push(name):
    nv = _lib()
    if nv is not None:
        nv.range_push("rvgp:" + name)


def pop():
    nv = _lib()
    if nv is not None:
        nv.range_pop()

"""Import the UNMODIFIED reference (``/root/reference/RVGP``) in this container.

TEST INFRASTRUCTURE ONLY — used by ``tests/golden/make_golden.py`` to generate golden
vectors and by ``tests/test_oracle_vs_reference.py`` (skipped when /root/reference is absent,
as on the GPU box).  Never imported by rvgp_b200/.

Accommodations (none edits a reference source file):
  1. ``RVGP/geometry.py:7`` imports tensorflow at module top but touches only ``tf.float64``
     (``:66``) and ``tf.convert_to_tensor`` (``:77-78``, consumed by ``.numpy()`` at
     ``dataclass.py:50,56-57``) -> a two-symbol stub module is registered as ``tensorflow``.
  2. ``RVGP/__init__.py`` imports kernels/main -> gpflow (absent).  A synthetic package object
     with ``__path__`` pointing at the reference is registered instead, and
     ``RVGP.geometry / RVGP.smoothing / RVGP.dataclass`` are imported directly.
  3. ``ptu_dijkstra`` (compiled unchanged into oracle/_ref by oracle/build_ref.py) takes
     ``int[:]`` CSR buffers (``pyx:221-229, 300-311``) but networkx 3.6 / SciPy 1.18 produce int64
     indices (the pinned networkx 3.1 / SciPy 1.10 gave int32).  ``networkx.adjacency_matrix`` is
     wrapped to cast indices/indptr to int32; values and structure are unchanged.
"""
import importlib
import os
import sys
import types

import numpy as np

REF_ROOT = "/root/reference"
_REF_SO_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "RVGP"))


class _FakeTensor:
    def __init__(self, a):
        self._a = np.asarray(a)

    def numpy(self):
        return self._a


def _install_tf_stub():
    if "tensorflow" in sys.modules:
        return
    tf = types.ModuleType("tensorflow")
    tf.float64 = np.float64
    tf.convert_to_tensor = lambda a, dtype=None: _FakeTensor(np.asarray(a, dtype=dtype))
    tf.__rvgp_stub__ = True
    sys.modules["tensorflow"] = tf


def _install_nx_int32():
    import networkx as nx
    if getattr(nx.adjacency_matrix, "__rvgp_int32__", False):
        return
    orig = nx.adjacency_matrix

    def adjacency_matrix_int32(G, nodelist=None, dtype=None, weight="weight"):
        from scipy import sparse
        M = sparse.csr_matrix(orig(G, nodelist=nodelist, dtype=dtype, weight=weight))
        M.indices = M.indices.astype(np.int32)
        M.indptr = M.indptr.astype(np.int32)
        return M

    adjacency_matrix_int32.__rvgp_int32__ = True
    nx.adjacency_matrix = adjacency_matrix_int32


def load(instrumented=False):
    """Returns a namespace with the reference's geometry, smoothing, dataclass modules and the
    compiled ptu_dijkstra module."""
    if not available():
        raise RuntimeError("reference tree not present (expected on the GPU box)")
    _install_tf_stub()
    _install_nx_int32()
    if _REF_SO_DIR not in sys.path:
        sys.path.insert(0, _REF_SO_DIR)
    ptu_name = "ptu_dijkstra_instr" if instrumented else "ptu_dijkstra"
    ptu = importlib.import_module(ptu_name)
    # dataclass.py:8 does `from ptu_dijkstra import connections, tangent_frames`
    sys.modules["ptu_dijkstra"] = ptu
    if "RVGP" not in sys.modules or not getattr(sys.modules["RVGP"], "__rvgp_ref__", False):
        for k in [k for k in sys.modules if k == "RVGP" or k.startswith("RVGP.")]:
            del sys.modules[k]
        pkg = types.ModuleType("RVGP")
        pkg.__path__ = [os.path.join(REF_ROOT, "RVGP")]
        pkg.__rvgp_ref__ = True
        sys.modules["RVGP"] = pkg
    ns = types.SimpleNamespace()
    ns.geometry = importlib.import_module("RVGP.geometry")
    ns.smoothing = importlib.import_module("RVGP.smoothing")
    ns.dataclass = importlib.import_module("RVGP.dataclass")
    ns.ptu = ptu
    return ns


def unload():
    """Remove the reference modules from sys.modules (so the repo's own RVGP shim can load)."""
    for k in [k for k in sys.modules if k == "RVGP" or k.startswith("RVGP.")]:
        if getattr(sys.modules.get("RVGP"), "__rvgp_ref__", False) or k != "RVGP":
            sys.modules.pop(k, None)
    sys.modules.pop("RVGP", None)
    sys.modules.pop("ptu_dijkstra", None)
    tf = sys.modules.get("tensorflow")
    if tf is not None and getattr(tf, "__rvgp_stub__", False):
        del sys.modules["tensorflow"]

"""Drop-in for the reference's only FFI module: the CPython extension ``ptu_dijkstra`` built from
RVGP/lib/ptu_dijkstra.pyx (setup.py:25-29) and imported at RVGP/dataclass.py:8.

Same call signatures (pyx:33, pyx:128), same exceptions (ValueError pyx:68-82,156-160; RuntimeError
pyx:119-123); the work runs in librvgp_b200.so (rvgp_geodesic_neighbourhoods, rvgp_tangent_frames,
rvgp_connections).  ``G`` may be a networkx graph (as in the reference) or an rvgp_b200 ManifoldGraph.
"""
import numpy as np
import torch

from rvgp_b200 import geometry as _geo


def tangent_frames(X, G, d, K):
    X = np.asarray(X)
    N, D = X.shape
    if K >= N:
        raise ValueError("Geodesic neighborhood size must be less than the total number of samples")
    if K < d:
        raise ValueError("Geodesic neighborhood size must be larger or equal to the embedding dimension")
    if D < d:
        raise ValueError("Embedding dimension must be less or equal to the ambient dimension of input data")
    g = _geo.ManifoldGraph.from_any(G, unit_weights_only=True)
    Xd = _geo.to_device_f64(X, g.indptr.device)
    seq, _ = _geo.geodesic_neighbourhoods_device(g.indptr, g.indices, int(K))
    T, S = _geo.tangent_frames_device(Xd, seq, int(d))
    return T.cpu().numpy(), S.cpu().numpy()


def connections(tangents, G, d):
    from scipy import sparse
    tangents = np.asarray(tangents)
    N, D = tangents.shape[0], tangents.shape[1]
    if D < d:
        raise ValueError("Embedding dimension must be less or equal to the ambient dimension of input data")
    g = _geo.ManifoldGraph.from_any(G)
    Td = _geo.to_device_f64(np.ascontiguousarray(tangents[:, :, :d]), g.indptr.device)
    _, R = _geo.connections_device(Td, g.indptr, g.indices, want_R=True)
    ip, ix = g.indptr.cpu().numpy(), g.indices.cpu().numpy()
    return sparse.bsr_matrix((R.cpu().numpy(), ix, ip), shape=(N * d, N * d)).tocoo()

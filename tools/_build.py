"""Shared set-up of the measurement scripts: the Morton-ordered connection Laplacian (BsrMatrix, d = 2) and scalar Laplacian (pattern
mode) of a synthetic cloud, built with the product's own kernels."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rvgp_b200 import geometry as geo
from rvgp_b200.eigensolver import BsrMatrix
from tests.workloads import make_cloud


def build(kind, n, nb=10):
    dev = torch.device("cuda", 0)
    Xd = torch.from_numpy(make_cloud(kind, n, 0)).to(dev)
    g = geo.manifold_graph(Xd, n_neighbors=nb)
    seq, _ = geo.geodesic_neighbourhoods_device(g.indptr, g.indices, int(nb * 1.5), g.max_row)
    T, S = geo.tangent_frames_device(Xd, seq, Xd.shape[1])
    gauges = geo.slice_frames_device(T, 2)
    order, inv = geo.morton_order_device(Xd)
    ip, ix = geo.csr_permute_device(g.indptr, g.indices, order, inv)
    gp = geo.gather_rows_device(gauges.reshape(n, -1), order).reshape(n, Xd.shape[1], 2)
    vals = geo.connections_device(gp, ip, ix)
    return BsrMatrix(n, 2, ip, ix, vals), BsrMatrix(n, 1, ip, ix, None), (g.indptr, g.indices, gauges)

"""Host-side mirror of the reference's ``RVGP/geometry.py`` over librvgp_b200.so.

Public names, argument meaning and error behaviour follow the reference:
``manifold_graph`` (geometry.py:100-123), ``compute_laplacian`` (:55-63),
``compute_connection_laplacian`` (:14-52), ``compute_spectrum`` (:66-80), ``manifold_dimension`` (:83-97),
``furthest_point_sampling`` (:126-162), ``project_to_manifold`` (:165-168), ``express_in_local_frame``
(:171-176).  Every numerical stage runs in hand-written sm_100a kernels through the C ABI; there is no CPU
path.  The ``*_device`` functions are the building blocks the data object uses (device tensors in / out).
"""
import math

import numpy as np
import torch

from ._cabi import get_handle, I64, RvgpError, RVGP_ERR_RANK_DEFICIENT
from .eigensolver import BsrMatrix, smallest_eigenpairs


def _dev(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("rvgp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def to_device_f64(x, device=None):
    """numpy / torch / DLPack producer -> contiguous float64 CUDA tensor."""
    dev = _dev(device)
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.float64).contiguous()
    if hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray):
        return torch.from_dlpack(x).to(device=dev, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(dev)


class _Tensor(np.ndarray):
    """ndarray with the ``.numpy()`` the reference calls on compute_spectrum's TF outputs (dataclass.py:50)."""

    def numpy(self):
        return np.asarray(self)


# ---------------------------------------------------------------------------------------------------------
# device building blocks
# ---------------------------------------------------------------------------------------------------------
def knn_device(Xd, n_neighbors, q_begin=0, q_count=None, return_d2=False, method="auto"):
    """K2.  (n, D) cuda f64 -> (q_count, nb) int32 neighbour ids, ascending (distance, index)."""
    h = get_handle(Xd.device.index)
    n, D = Xd.shape
    if q_count is None:
        q_count = n - q_begin
    if n_neighbors >= n:
        raise ValueError("Expected n_neighbors < n_samples_fit, but n_neighbors = %d, n_samples_fit = %d"
                         % (n_neighbors, n))
    idx = torch.empty((q_count, n_neighbors), dtype=torch.int32, device=Xd.device)
    d2 = torch.empty((q_count, n_neighbors), dtype=torch.float64, device=Xd.device) if return_d2 else None
    if method == "auto":
        method = "grid" if (D <= 3 and n >= 4096) else "brute"
    if method == "grid":
        import ctypes
        mm = torch.stack([Xd.min(0).values, Xd.max(0).values]).cpu().numpy()       # bounding box (plumbing)
        lo = (ctypes.c_double * D)(*mm[0].tolist())
        hi = (ctypes.c_double * D)(*mm[1].tolist())
        wsb = h.query("rvgp_knn_grid_workspace_bytes", int(n), int(D), lo, hi)
        ws = torch.empty(wsb, dtype=torch.uint8, device=Xd.device)
        h.call("rvgp_knn_grid_f64", Xd, int(n), int(D), lo, hi, int(q_begin), int(q_count), int(n_neighbors), idx, d2,
               ws, I64(wsb))
    else:
        h.call("rvgp_knn_f64", Xd, int(n), int(D), int(q_begin), int(q_count), int(n_neighbors), idx, d2)
    return (idx, d2) if return_d2 else idx


def knn_to_csr_device(knn):
    """K3.  Directed kNN lists -> symmetric CSR with self loops (union), int32, sorted columns."""
    h = get_handle(knn.device.index)
    n, k = knn.shape
    wsb = h.query("rvgp_knn_to_csr_workspace_bytes", int(n), int(k))
    ws = torch.empty(wsb, dtype=torch.uint8, device=knn.device)
    indptr = torch.empty(n + 1, dtype=torch.int32, device=knn.device)
    indices = torch.empty(2 * n * k + n, dtype=torch.int32, device=knn.device)
    nnz = torch.zeros(1, dtype=torch.int32, device=knn.device)
    h.call("rvgp_knn_to_csr", knn, int(n), int(k), indptr, indices, nnz, ws, I64(wsb))
    nnz = int(nnz.item())
    return indptr, indices[:nnz].clone()


def geodesic_neighbourhoods_device(indptr, indices, K, maxrow=None, comm=None):
    """K4.  Returns seq (n, K+1) int32 in the reference heap's pop order, counts (n).  ``comm`` (distributed.Comm, world > 1):
    the sources are independent, so every rank runs a contiguous range and the rows are all-gathered."""
    h = get_handle(indptr.device.index)
    n = indptr.numel() - 1
    K = int(K)
    if K >= n:
        raise ValueError("Geodesic neighborhood size must be less than the total number of samples")
    if maxrow is None:
        maxrow = int((indptr[1:] - indptr[:-1]).max().item())
    wsb = h.query("rvgp_geodesic_workspace_bytes", h._h, int(n), K, int(maxrow))
    ws = torch.empty(wsb, dtype=torch.uint8, device=indptr.device)
    seq = torch.zeros((n, K + 1), dtype=torch.int32, device=indptr.device)
    counts = torch.empty(n, dtype=torch.int32, device=indptr.device)
    flags = torch.zeros(1, dtype=torch.int32, device=indptr.device)
    if comm is not None and comm.world > 1:
        s0, s1 = (n * comm.rank) // comm.world, (n * (comm.rank + 1)) // comm.world
        h.call("rvgp_geodesic_neighbourhoods_range", indptr, indices, int(n), K, int(maxrow), int(s0), int(s1 - s0), seq, counts,
               flags, ws, I64(wsb))
        rows = [(n * (r + 1)) // comm.world - (n * r) // comm.world for r in range(comm.world)]
        both = torch.cat([seq[s0:s1], counts[s0:s1, None]], dim=1)
        both = comm.allgather_rows(both, rows)
        seq, counts = both[:, :K + 1].contiguous(), both[:, K + 1].contiguous()
        bits = torch.stack([(flags >> i) & 1 for i in range(3)]).reshape(3).to(torch.int32)      # OR over the ranks, bit by bit
        comm.allreduce_(bits)
        flags = ((bits[0] > 0).to(torch.int32) | ((bits[1] > 0).to(torch.int32) << 1) | ((bits[2] > 0).to(torch.int32) << 2)).reshape(1)
        h.call("rvgp_geodesic_fix_stale", int(n), K, seq, counts, flags)
    else:
        h.call("rvgp_geodesic_neighbourhoods", indptr, indices, int(n), K, int(maxrow), seq, counts, flags, ws, I64(wsb))
    f = int(flags.item())
    if f & 2:
        raise RvgpError(-1, "geodesic: decrease_val would fire (non-unit weights are not supported)")
    if f & 4:
        raise RvgpError(-6, "geodesic: node pool overflow")
    return seq, counts


def tangent_frames_device(Xd, seq, d):
    """K5.  tangents (n, D, d) (first d left singular vectors), Sigma (n, d).  Raises RuntimeError like
    ptu_dijkstra.pyx:119-123 when a neighbourhood does not span d dimensions."""
    h = get_handle(Xd.device.index)
    n, D = Xd.shape
    Kp1 = seq.shape[1]
    T = torch.empty((n, D, D), dtype=torch.float64, device=Xd.device)
    S = torch.empty((n, D), dtype=torch.float64, device=Xd.device)
    flag = torch.zeros(1, dtype=torch.int32, device=Xd.device)
    h.call("rvgp_tangent_frames", Xd, int(n), int(D), seq, int(Kp1), int(d), T, S, flag)
    if int(flag.item()) & 1:
        raise RuntimeError("Local tangent space approximation failed, at least one geodesic "
                           "neighborhood does not span d-dimensional space")
    if d < D:
        return slice_frames_device(T, d), S[:, :d].contiguous()
    return T, S


def slice_frames_device(T, d):
    h = get_handle(T.device.index)
    n, D, dfull = T.shape
    if d == dfull:
        return T
    G = torch.empty((n, D, d), dtype=torch.float64, device=T.device)
    h.call("rvgp_slice_frames", T, I64(n), int(D), int(dfull), int(d), G)
    return G


def explained_variance_device(Sigma):
    """K6.  mean - std (population) over nodes of the cumulative explained variance (geometry.py:89-92)."""
    h = get_handle(Sigma.device.index)
    n, D = Sigma.shape
    cum = torch.empty_like(Sigma)
    h.call("rvgp_sigma_cumvar", Sigma, int(n), int(D), cum)
    ws = torch.empty(max(1, h.query("rvgp_coldot_workspace_bytes", I64(n), int(D)) // 8), dtype=torch.float64,
                     device=Sigma.device)
    s = torch.empty(D, dtype=torch.float64, device=Sigma.device)
    h.call("rvgp_coldot_f64", I64(n), int(D), cum, I64(D), None, I64(0), s, ws)
    mean = s / n
    h.call("rvgp_resid_sq_f64", I64(n), int(D), cum, I64(D), None, I64(0), mean, s, ws)
    var_exp = mean.cpu().numpy() - np.sqrt(s.cpu().numpy() / n)
    return var_exp


def connections_device(gauges, indptr, indices, want_R=False):
    """K7+K8.  Returns Lc block values (nnzb, d, d) [and the raw R blocks]."""
    h = get_handle(gauges.device.index)
    n, D, d = gauges.shape
    nnzb = indices.numel()
    Lc = torch.empty((nnzb, d, d), dtype=torch.float64, device=gauges.device)
    R = torch.empty((nnzb, d, d), dtype=torch.float64, device=gauges.device) if want_R else None
    h.call("rvgp_connections", gauges, int(n), int(D), int(d), indptr, indices, I64(nnzb), Lc, R)
    return (Lc, R) if want_R else Lc


def orient_gauges_device(indptr, indices, Lc_vals, rounds=64, steps_per_round=128):
    """K20.  Node signs s_i = +-1 (int32 cuda) with s_i s_j det(block(i, j)) = +1 on every stored edge of the d = 2
    connection Laplacian, or None when no such signs exist (non-orientable surface, or a noisy tangent plane that breaks a
    cycle).  Label propagation over the block-CSR pattern, one seed per connected component; verified edge by edge."""
    h = get_handle(indptr.device.index)
    n = int(indptr.numel()) - 1
    dev = indptr.device
    labels = torch.zeros(n, dtype=torch.int32, device=dev)
    changed = torch.zeros(1, dtype=torch.int32, device=dev)
    out2 = torch.zeros(2, dtype=torch.int32, device=dev)
    labels[0] = 1
    for _ in range(rounds * 64):
        changed.zero_()
        h.call("rvgp_orient_steps", int(n), indptr, indices, Lc_vals, labels, changed, int(steps_per_round))
        if int(changed.item()) != 0:
            continue
        h.call("rvgp_orient_check", int(n), indptr, indices, Lc_vals, labels, out2)
        bad, unlabelled = [int(v) for v in out2.cpu().tolist()]
        if unlabelled == 0:
            return labels if bad == 0 else None
        # another connected component: seed its first node
        first = int(torch.nonzero(labels == 0)[0].item())
        labels[first] = 1
    return None


def morton_order_device(Xd):
    h = get_handle(Xd.device.index)
    n, D = Xd.shape
    wsb = h.query("rvgp_morton_order_workspace_bytes", int(n))
    ws = torch.empty(wsb, dtype=torch.uint8, device=Xd.device)
    order = torch.empty(n, dtype=torch.int32, device=Xd.device)
    inv = torch.empty(n, dtype=torch.int32, device=Xd.device)
    h.call("rvgp_morton_order", Xd, int(n), int(D), order, inv, ws, I64(wsb))
    return order, inv


def csr_permute_device(indptr, indices, order, inv):
    h = get_handle(indptr.device.index)
    n = indptr.numel() - 1
    wsb = h.query("rvgp_csr_permute_workspace_bytes", int(n))
    ws = torch.empty(wsb, dtype=torch.uint8, device=indptr.device)
    ip = torch.empty_like(indptr)
    ix = torch.empty_like(indices)
    h.call("rvgp_csr_permute", int(n), indptr, indices, order, inv, ip, ix, ws, I64(wsb))
    return ip, ix


def gather_rows_device(A2d, perm, block=1):
    """out[r] = A[perm[r // block] * block + r % block]."""
    h = get_handle(A2d.device.index)
    nrows, ncols = A2d.shape
    out = torch.empty_like(A2d)
    h.call("rvgp_gather_rows_f64", I64(nrows), int(ncols), A2d, I64(A2d.stride(0)), perm, int(block), out,
           I64(out.stride(0)))
    return out


def frame_apply_device(gauges, x, mode, scale=1.0):
    """mode 0: G^T x  (n,D,nc)->(n,d,nc);  mode 1: G x  (n,d,nc)->(n,D,nc)."""
    h = get_handle(gauges.device.index)
    n, D, d = gauges.shape
    x3 = x.reshape(n, -1, 1) if x.dim() == 2 else x
    nc = x3.shape[2]
    out = torch.empty((n, d if mode == 0 else D, nc), dtype=torch.float64, device=gauges.device)
    h.call("rvgp_frame_apply", gauges, I64(n), int(D), int(d), x3.contiguous(), out, int(nc), int(mode), float(scale))
    return out.reshape(n, -1) if x.dim() == 2 else out


# ---------------------------------------------------------------------------------------------------------
# public API (reference names)
# ---------------------------------------------------------------------------------------------------------
class ManifoldGraph:
    """Device CSR of a symmetric graph (the symmetrised kNN graph with self loops on the hot path).  ``weights`` is None
    for unit weights (the hot path: pattern only) or a device f64 tensor, one weight per stored entry (typ='affinity',
    weighted networkx input).  ``to_networkx()`` materialises the reference's networkx object (geometry.py:112,120-121)
    on demand."""

    def __init__(self, indptr, indices, X=None, knn=None, weights=None, dense=False):
        self.indptr, self.indices, self.X, self.knn, self.weights = indptr, indices, X, knn, weights
        self.dense = bool(dense)          # every (i, j) pair stored (typ='affinity')
        self.n = indptr.numel() - 1
        self._nx = None

    def __len__(self):
        return self.n

    def number_of_nodes(self):
        return self.n

    @property
    def max_row(self):
        return int((self.indptr[1:] - self.indptr[:-1]).max().item())

    def scipy_adjacency(self):
        from scipy import sparse
        ip = self.indptr.cpu().numpy()
        ix = self.indices.cpu().numpy()
        w = np.ones(ix.size) if self.weights is None else self.weights.cpu().numpy().reshape(-1)
        return sparse.csr_matrix((w, ix, ip), shape=(self.n, self.n))

    def to_networkx(self):
        if self._nx is None:
            import networkx as nx
            if self.dense:                # nx.from_numpy_array(A) as in geometry.py:118 (zero weights are no edges)
                G = nx.from_numpy_array(self.weights.reshape(self.n, self.n).cpu().numpy())
            else:
                G = nx.from_scipy_sparse_array(self.scipy_adjacency())
            if self.X is not None:
                Xh = self.X.cpu().numpy() if isinstance(self.X, torch.Tensor) else np.asarray(self.X)
                nx.set_node_attributes(G, {i: Xh[i] for i in G.nodes}, "pos")
            self._nx = G
        return self._nx

    @staticmethod
    def from_any(G, device=None, unit_weights_only=False):
        """Accept a ManifoldGraph or a networkx graph (the reference's inner-FFI argument type): adjacency with weights,
        symmetrised with the element-wise maximum like pyx:84-103.  ``unit_weights_only``: the caller's kernel walks the
        pattern with unit edge lengths (the heap-Dijkstra emulation) -- anything else is refused, not silently changed."""
        if not isinstance(G, ManifoldGraph):
            import networkx as nx
            from scipy import sparse
            M = sparse.csr_matrix(nx.adjacency_matrix(G, weight="weight"), dtype=np.float64)
            M = M.maximum(M.T).tocsr()
            M.sort_indices()
            dev = _dev(device)
            unit = bool(np.all(M.data == 1.0))
            G = ManifoldGraph(torch.from_numpy(M.indptr.astype(np.int32)).to(dev),
                              torch.from_numpy(M.indices.astype(np.int32)).to(dev),
                              weights=None if unit else torch.from_numpy(np.ascontiguousarray(M.data)).to(dev))
        if unit_weights_only and G.weights is not None:
            raise NotImplementedError("geodesic neighbourhoods on a graph with non-unit edge weights are not on the B200 "
                                      "hot path (the reference pipeline only builds unit-weight kNN graphs, geometry.py:103-112)")
        return G


def manifold_graph(X, typ="knn", n_neighbors=5, device=None):
    """Fit graph over a pointset X (geometry.py:100-123).

    typ='knn' (the hot path): directed kNN -> + I -> undirected union, unit weights.  Limits of the CUDA kernels:
    n_neighbors <= 32 and ambient dimension D <= 64 (sklearn, which the reference calls, has no such limits).
    typ='affinity' (geometry.py:114-118): the DENSE Gaussian-kernel graph exp(-dist^2 / (2 * 0.1^2)) on all pairs incl. the
    diagonal (weight 1 self loops); n x n weights on the device, n <= 65535.  Its Laplacians come from compute_laplacian;
    the unit-weight heap-Dijkstra stage (tangent_frames) refuses weighted graphs."""
    Xd = to_device_f64(X, device)
    if typ == "knn":
        knn = knn_device(Xd, n_neighbors)
        indptr, indices = knn_to_csr_device(knn)
        return ManifoldGraph(indptr, indices, X=Xd, knn=knn)
    if typ == "affinity":
        n, D = Xd.shape
        h = get_handle(Xd.device.index)
        A = torch.empty((n, n), dtype=torch.float64, device=Xd.device)
        ws = torch.empty(max(1, n), dtype=torch.float64, device=Xd.device)
        sigma = 0.1                                        # "Control the width of the Gaussian kernel" (geometry.py:116)
        h.call("rvgp_affinity_f64", Xd, int(n), int(D), float(sigma), A, ws)
        indptr = (torch.arange(n + 1, device=Xd.device, dtype=torch.int64) * n).to(torch.int32)
        indices = torch.arange(n, device=Xd.device, dtype=torch.int32).repeat(n)
        return ManifoldGraph(indptr, indices, X=Xd, weights=A.reshape(-1), dense=True)
    raise ValueError("manifold_graph: typ must be 'knn' or 'affinity', got %r" % (typ,))   # the reference hits a NameError here


def manifold_dimension(Sigma, frac_explained=0.9):
    """Estimate manifold dimension based on singular values (geometry.py:83-97)."""
    Sd = to_device_f64(Sigma)
    if frac_explained == 1.0:
        return Sd.shape[1]
    var_exp = explained_variance_device(Sd)
    dim_man = np.where(var_exp >= frac_explained)[0][0] + 1
    print("Fraction of variance explained: ", var_exp)
    return int(dim_man)


def _adjacency_host(G):
    """Host scipy CSR adjacency WITH weights, the matrix networkx hands to its Laplacian builders
    (``nx.to_scipy_sparse_array(G, weight='weight')``: a self loop appears once on the diagonal).  From a ManifoldGraph
    (device -> host copy; unit weights unless the graph carries some) or straight from a networkx graph (the reference's
    argument type; no device involved)."""
    from scipy import sparse
    if isinstance(G, ManifoldGraph):
        return G.scipy_adjacency()
    import networkx as nx
    A = sparse.csr_matrix(nx.to_scipy_sparse_array(G, weight="weight", format="csr"), dtype=np.float64)
    A.sort_indices()
    return A


def compute_laplacian(G, normalization=False):
    """Graph Laplacian as scipy CSR f64 (geometry.py:55-63) = networkx's ``laplacian_matrix`` / ``normalized_laplacian_matrix``
    restated on the host: L = diag(rowsum(A)) - A with an explicit diagonal, edge weights kept (so loop-free and weighted
    graphs, e.g. typ='affinity', are handled like the pipeline's unit-weight kNN graph with self loops, where the self
    loop cancels and the diagonal is the number of non-self neighbours); ``normalization=True``: D^-1/2 L D^-1/2 with
    D = rowsum(A) INCLUDING the self loop, 1/sqrt(0) -> 0."""
    from scipy import sparse
    A = _adjacency_host(G)
    n = A.shape[0]
    dsum = np.asarray(A.sum(axis=1)).reshape(-1)
    L = sparse.csr_matrix(sparse.diags(dsum, 0, shape=(n, n), format="csr") - A, dtype=np.float64)
    if normalization:
        with np.errstate(divide="ignore"):
            dh = 1.0 / np.sqrt(dsum)
        dh[np.isinf(dh)] = 0.0
        DH = sparse.diags(dh, 0, shape=(n, n), format="csr")
        L = sparse.csr_matrix(DH @ (L @ DH), dtype=np.float64)
    L.sort_indices()
    return L


def _nx_degree(G, A=None):
    """networkx's unweighted ``G.degree()``: number of incident edges, a self loop counted twice (geometry.py:46)."""
    if not isinstance(G, ManifoldGraph):
        return np.array(list(dict(G.degree()).values()), dtype=np.float64)
    A = G.scipy_adjacency() if A is None else A
    n = A.shape[0]
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    has_loop = np.zeros(n, dtype=bool)
    has_loop[rows[rows == A.indices]] = True
    return (np.diff(A.indptr) + has_loop).astype(np.float64)


def compute_connection_laplacian(G, R, normalization=None):
    """Connection Laplacian as scipy sparse (geometry.py:14-52): kron(L, 1_{dxd}) .* R; ``normalization='rw'`` multiplies
    row block i by 1 / deg_i with networkx's degree (a self loop counts twice, geometry.py:45-50)."""
    from scipy import sparse
    n = len(G)
    dim = R.shape[0] // n
    L = compute_laplacian(G)
    Lc = sparse.kron(L, np.ones([dim, dim])).multiply(R)
    if normalization == "rw":
        deg = _nx_degree(G)
        with np.errstate(divide="ignore"):
            deg_inv = 1.0 / deg
        deg_inv[np.isinf(deg_inv)] = 0
        return sparse.diags(deg_inv.repeat(dim), 0, format="csr") @ Lc
    return sparse.bsr_matrix(Lc, blocksize=(dim, dim))


def compute_spectrum(laplacian, n_eigenpairs=None, dtype=None, tol=1e-12):
    """Smallest-k eigenpairs (geometry.py:66-80): ascending eigenvalues, eigenvectors scaled by sqrt(#rows).
    Accepts a scipy sparse matrix (CSR or BSR) and runs the GPU block eigensolver."""
    from scipy import sparse
    N = laplacian.shape[0]
    if n_eigenpairs is None or n_eigenpairs >= N:
        n_eigenpairs = N
    dev = _dev()
    if sparse.isspmatrix_bsr(laplacian) and laplacian.blocksize[0] == laplacian.blocksize[1]:
        d = laplacian.blocksize[0]
        B = laplacian
    else:
        d = 1
        B = sparse.csr_matrix(laplacian)
        B.sort_indices()
    vals = torch.from_numpy(np.ascontiguousarray(B.data, dtype=np.float64).reshape(-1, d, d)).to(dev)
    A = BsrMatrix(N // d, d, torch.from_numpy(B.indptr.astype(np.int32)).to(dev),
                  torch.from_numpy(B.indices.astype(np.int32)).to(dev), vals)
    # Gershgorin bound for symmetric matrices: max absolute row sum
    hi = float(abs(sparse.csr_matrix(laplacian)).sum(1).max())
    st = {}
    evals, evecs = smallest_eigenpairs(A, n_eigenpairs, upper_bound=hi, lower_bound=min(0.0, -1e-12 * hi), tol=tol, stats=st)
    if not st.get("converged", True):
        # scipy's eigsh raises ArpackNoConvergence here (geometry.py:73)
        raise RuntimeError("compute_spectrum: eigensolver did not converge (max residual %.3e > %.3e)" %
                           (st["residual_max"], st["tol_abs"]))
    evecs = evecs.cpu().numpy() * np.sqrt(N)
    return evals.cpu().numpy().view(_Tensor), evecs.view(_Tensor)


def project_to_manifold(x, gauges):
    """Project vectors onto the local tangent frames (geometry.py:165-168)."""
    Gd = to_device_f64(gauges)
    coeffs = frame_apply_device(Gd, to_device_f64(x), 0)
    return frame_apply_device(Gd, coeffs, 1).cpu().numpy()


def express_in_local_frame(x, gauges, reverse=False):
    """Express vectors in local coordinates (geometry.py:171-176)."""
    Gd = to_device_f64(gauges)
    return frame_apply_device(Gd, to_device_f64(x), 1 if reverse else 0).cpu().numpy()


def furthest_point_sampling(x, N=None, spacing=0.1, start_idx=0, stop_crit=None):
    """Greedy furthest-point sampling (geometry.py:126-162).  ``stop_crit`` is the README's stale alias of
    ``spacing`` (README.md:70)."""
    from .fps import furthest_point_sampling_device
    if stop_crit is not None:
        spacing = stop_crit
    if spacing == 0.0:
        return np.arange(len(x)), None
    perm, lambdas = furthest_point_sampling_device(to_device_f64(x), N, spacing, start_idx)
    return perm.cpu().numpy(), lambdas.cpu().numpy()

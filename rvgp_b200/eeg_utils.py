"""Mirror of the reference's ``examples/eeg_example/eeg_utils.py`` pieces that sit next to the hot path:
``compute_vectorfield_features`` (div / curl from k nearest neighbours, :46-80), its time loop (:36-44),
``project_to_local_frame`` (:19-23), ``interpolate_time_range`` / ``interpolate_timepoint`` (:26-35, :83-116) and the error
measures (:119-131).  kNN and the feature sums run on the device (K2 + K19); the per-frame fits reuse the fixed eigenbasis
and the fused small-k GP kernel (K16)."""
import numpy as np
import torch

from ._cabi import get_handle
from .geometry import express_in_local_frame, knn_device, to_device_f64


def project_to_local_frame(x, gauges, reverse=False):
    return express_in_local_frame(x, gauges, reverse=reverse)


def compute_vectorfield_features(positions, vectors, k=5, reference_row0=True, knn=None):
    """(div (n,), curl (n, 3)).  ``reference_row0=True`` reproduces eeg_utils.py:67, which subtracts
    ``normalized_vectors[0]`` (row 0) instead of the point's own vector; pass False for the intended estimator.
    ``knn`` (n, k) int32 cuda: reuse neighbour lists across time frames."""
    P = to_device_f64(positions)
    V = to_device_f64(vectors, P.device)
    n = P.shape[0]
    if P.shape[1] != 3 or V.shape != P.shape:
        raise ValueError("positions and vectors must both be (n, 3)")
    if knn is None:
        knn = knn_device(P, int(k))
    h = get_handle(P.device.index)
    div = torch.empty(n, dtype=torch.float64, device=P.device)
    curl = torch.empty((n, 3), dtype=torch.float64, device=P.device)
    h.call("rvgp_vectorfield_features_f64", int(n), int(knn.shape[1]), P, V, knn, int(bool(reference_row0)), div, curl)
    return div.cpu().numpy(), curl.cpu().numpy()


def compute_vectorfield_features_time(timepoints, positions, vectors, k=5, reference_row0=True):
    """eeg_utils.py:36-44; the neighbour lists are computed once (positions do not change over time)."""
    P = to_device_f64(positions)
    knn = knn_device(P, int(k))
    div = np.zeros([len(timepoints), positions.shape[0]])
    curl = np.zeros([len(timepoints), positions.shape[0], 3])
    for t in timepoints:
        div[t, :], curl[t, :, :] = compute_vectorfield_features(P, vectors[t, :, :], k=k, reference_row0=reference_row0, knn=knn)
    return div, curl


def interpolate_timepoint(d, train_idx, test_idx, project=True, t=0, plot=False, dim_emb=3, dim_man=2, n_eigenpairs=50):
    """eeg_utils.py:83-116.  ``project``: the signal at the training nodes is expressed in the local frames, optionally diffused
    for time ``t`` on the SUB-GRAPH of the training nodes (principal sub-matrices of Lc and L, :99-104), and mapped back.  The
    sub-matrices are sliced sparsely (the reference goes through the dense ``d.Lc.A``) and the diffusion is the Chebyshev
    exp(-tA) action of smoothing.vector_diffusion(method="matrix_exp") on the device."""
    import RVGP
    if project:
        # The reference replaces d.vectors by the (n_train, D) result, which its own train_gp then indexes with node ids
        # (main.py:25) -> IndexError unless the training nodes are 0..n_train-1.  Here the processed rows are written back
        # into a full (n, D) field (identical whenever the reference works; defined in all other cases).
        train_idx = np.asarray(train_idx)
        full = np.array(d.vectors, dtype=np.float64, copy=True)
        if full.shape[0] == len(train_idx) and full.shape[0] != d.n:        # caller passed the training rows only
            tmp = np.zeros((d.n, full.shape[1]))
            tmp[train_idx] = full
            full = tmp
        v = project_to_local_frame(full[train_idx], d.gauges[train_idx, :, :])
        if t > 0:
            from scipy import sparse
            from .geometry import compute_laplacian
            from .smoothing import vector_diffusion
            dm = d.gauges.shape[2]
            Lc_idx = np.sort(np.hstack([train_idx * dm + q for q in range(dm)]))     # the reference hard-codes dm = 2 (:100)
            Lc_ = sparse.bsr_matrix(d.Lc.tocsr()[Lc_idx, :][:, Lc_idx], blocksize=(dm, dm))
            L = compute_laplacian(d._graph if hasattr(d, "_graph") else d.G)
            L = L[train_idx, :][:, train_idx]
            v = vector_diffusion(v, t, L=L, Lc=Lc_, method="matrix_exp")
        full[train_idx] = project_to_local_frame(v, d.gauges[train_idx, :, :], reverse=True)
        d.vectors = full
    gp = RVGP.fit(d, train_ind=train_idx, epochs=100, noise_variance=0.001)
    f_pred, _ = gp.transform(d, test_idx)
    return f_pred


def interpolate_time_range(timepoints, X, f, train_idx, test_idx):
    """eeg_utils.py:26-35: one eigenbasis (n_eigenpairs=10), one GP fit + transform per time frame."""
    import RVGP
    d = RVGP.create_data_object(X, n_eigenpairs=10)
    f_pred = np.zeros([len(timepoints), X.shape[0], X.shape[1]])        # like the reference: test_idx must cover every node
    for t, time in enumerate(timepoints):
        d.vectors = f[int(time), :, :]
        f_pred[t, :, :] = interpolate_timepoint(d, train_idx, test_idx, project=False)
    return f_pred


def compute_error(f_pred, f_real, test_idx, plot=False, verbose=False):
    l2_error = np.linalg.norm(f_real[test_idx, :].ravel() - f_pred[test_idx, :].ravel()) / len(f_real[test_idx, :].ravel())
    if verbose:
        print("Relative l2 error is {}".format(l2_error))
    return l2_error


def compute_error_time(f_pred, f_real, test_idx):
    squared_diffs = (f_real - f_pred) ** 2
    sum_squared_diffs = np.einsum('ijk->i', squared_diffs)
    return np.sqrt(sum_squared_diffs)

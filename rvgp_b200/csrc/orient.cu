// K20: consistent orientation of the d = 2 gauges, so that the connection Laplacian becomes a complex-Hermitian operator.
//
// The reference's connection blocks R_ij = U V^T of SVD(T_i^T T_j) (pyx:259-293) are orthogonal 2x2 matrices whose
// determinant is +1 or -1 depending on whether the two local PCA frames happen to have the same handedness (the sign of
// a singular vector is arbitrary).  On an orientable surface there are node signs s_i = +-1 with s_i s_j det(R_ij) = +1 for
// every edge: flipping the second gauge vector of the nodes with s_i = -1 is a similarity transform D Lc D, D = diag(1, s_i),
// that turns EVERY block into a scaled rotation [[a, -b], [b, a]].  Such a matrix commutes with the per-node quarter turn J,
// i.e. it is an n x n complex-Hermitian matrix acting on z_i = x_i + i y_i; each of its eigenvalues is an exactly double
// eigenvalue of Lc (the pairs the reference's ARPACK output shows), and the block eigensolver needs HALF the
// columns (rvgp_b200/eigensolver.py: smallest_eigenpairs_paired).
//
//   orient_step_kernel   pull-style label propagation: an unlabelled node takes s_i = s_j * sign(det block(i,j)) from any
//                        labelled neighbour (benign race: labels are written once; inconsistency is caught by the check)
//   orient_check_kernel  counts edges with s_i s_j det < 0 and unlabelled nodes
//   rot90_nodes_kernel   J: out[2i] = -V[2i+1], out[2i+1] = V[2i]   (multiplication by i in the complex picture)
#include "common.cuh"

namespace rvgp {

__device__ __forceinline__ int det_sign(const double* __restrict__ b) {
    return (b[0] * b[3] - b[1] * b[2]) < 0.0 ? -1 : 1;
}

__global__ void __launch_bounds__(256)
orient_step_kernel(int n, const int* __restrict__ indptr, const int* __restrict__ indices, const double* __restrict__ vals,
                   int* labels, int* __restrict__ changed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (*reinterpret_cast<volatile int*>(labels + i) != 0) return;
    for (int e = indptr[i]; e < indptr[i + 1]; ++e) {
        const int j = indices[e];
        if (j == i) continue;
        const int sj = *reinterpret_cast<volatile int*>(labels + j);
        if (sj != 0) {
            labels[i] = sj * det_sign(vals + (int64_t)e * 4);
            *changed = 1;
            return;
        }
    }
}

// out[0] += number of stored off-diagonal blocks whose orientation disagrees with the labels; out[1] += unlabelled nodes
__global__ void __launch_bounds__(256)
orient_check_kernel(int n, const int* __restrict__ indptr, const int* __restrict__ indices, const double* __restrict__ vals,
                    const int* __restrict__ labels, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int si = labels[i];
    if (si == 0) { atomicAdd(out + 1, 1); return; }
    int bad = 0;
    for (int e = indptr[i]; e < indptr[i + 1]; ++e) {
        const int j = indices[e];
        if (j == i) continue;
        if (si * labels[j] * det_sign(vals + (int64_t)e * 4) < 0) ++bad;
    }
    if (bad) atomicAdd(out, bad);
}

__global__ void rot90_nodes_kernel(int64_t nnodes, int ncols, const double* __restrict__ V, int64_t ldv, double* __restrict__ out,
                                   int64_t ldo) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nnodes * ncols) return;
    const int64_t i = idx / ncols;
    const int c = (int)(idx % ncols);
    const double x = V[(2 * i) * ldv + c], y = V[(2 * i + 1) * ldv + c];
    out[(2 * i) * ldo + c] = -y;
    out[(2 * i + 1) * ldo + c] = x;
}

// Paired (complex-Hermitian) block algebra on real storage: a complex block W is one real (2 n x b) panel, i W = J W.
// pair_panel:   out (2 n x 2 b) = [V | J V]: ONE tall-skinny Gram  C^T [W | J W]  then yields Re and -Im of  C^H W  together.
// pair_combine: W = beta W + alpha (T[:, :b] + J T[:, b:2b]): ONE product  T = V [Re C | Im C]  then gives  V C = V Re C + J (V Im C).
__global__ void pair_panel_kernel(int64_t nnodes, int ncols, const double* __restrict__ V, int64_t ldv, double* __restrict__ out,
                                  int64_t ldo) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nnodes * ncols) return;
    const int64_t i = idx / ncols;
    const int c = (int)(idx % ncols);
    const double x = V[(2 * i) * ldv + c], y = V[(2 * i + 1) * ldv + c];
    out[(2 * i) * ldo + c] = x;
    out[(2 * i + 1) * ldo + c] = y;
    out[(2 * i) * ldo + ncols + c] = -y;
    out[(2 * i + 1) * ldo + ncols + c] = x;
}

__global__ void pair_combine_kernel(int64_t nnodes, int ncols, const double* __restrict__ T, int64_t ldt, double alpha, double beta,
                                    double* __restrict__ W, int64_t ldw) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nnodes * ncols) return;
    const int64_t i = idx / ncols;
    const int c = (int)(idx % ncols);
    const double rx = T[(2 * i) * ldt + c], ry = T[(2 * i + 1) * ldt + c];
    const double ix = T[(2 * i) * ldt + ncols + c], iy = T[(2 * i + 1) * ldt + ncols + c];
    double* w0 = W + (2 * i) * ldw + c;
    double* w1 = W + (2 * i + 1) * ldw + c;
    const double v0 = alpha * (rx - iy), v1 = alpha * (ry + ix);        // (J t)[2i] = -t[2i+1], (J t)[2i+1] = t[2i]
    *w0 = (beta != 0.0) ? fma(beta, *w0, v0) : v0;
    *w1 = (beta != 0.0) ? fma(beta, *w1, v1) : v1;
}

// V[2i+1, :] *= s[i]   (the similarity transform D = diag(1, s_i) applied to block vectors / gauges' second coordinate)
__global__ void flip_odd_rows_kernel(int64_t nnodes, int ncols, const int* __restrict__ s, double* __restrict__ V, int64_t ldv) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nnodes * ncols) return;
    const int64_t i = idx / ncols;
    const int c = (int)(idx % ncols);
    if (s[i] < 0) V[(2 * i + 1) * ldv + c] = -V[(2 * i + 1) * ldv + c];
}

}  // namespace rvgp

using namespace rvgp;

extern "C" int rvgp_orient_steps(rvgp_handle_t hh, int n, const int32_t* indptr, const int32_t* indices, const double* vals,
                                 int32_t* labels, int32_t* changed, int steps) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, n >= 1 && steps >= 1 && indptr && indices && vals && labels && changed, "orient_steps: bad arguments");
    for (int s = 0; s < steps; ++s) {
        orient_step_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(n, indptr, indices, vals, labels, changed);
        RVGP_LAUNCH_OK(h, "orient_step_kernel");
    }
    return RVGP_OK;
}

extern "C" int rvgp_orient_check(rvgp_handle_t hh, int n, const int32_t* indptr, const int32_t* indices, const double* vals,
                                 const int32_t* labels, int32_t* out2) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, n >= 1 && indptr && indices && vals && labels && out2, "orient_check: bad arguments");
    RVGP_CUDA_OK(h, cudaMemsetAsync(out2, 0, 2 * sizeof(int), h->stream));
    orient_check_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(n, indptr, indices, vals, labels, out2);
    RVGP_LAUNCH_OK(h, "orient_check_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_rot90_nodes_f64(rvgp_handle_t hh, int64_t nnodes, int ncols, const double* V, int64_t ldv, double* out,
                                    int64_t ldo) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, nnodes >= 0 && ncols >= 0 && V != out, "rot90_nodes: bad arguments (out must not alias V)");
    if (nnodes * ncols == 0) return RVGP_OK;
    rot90_nodes_kernel<<<cdiv(nnodes * ncols, 256), 256, 0, h->stream>>>(nnodes, ncols, V, ldv, out, ldo);
    RVGP_LAUNCH_OK(h, "rot90_nodes_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_pair_panel_f64(rvgp_handle_t hh, int64_t nnodes, int ncols, const double* V, int64_t ldv, double* out,
                                   int64_t ldo) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, nnodes >= 0 && ncols >= 0 && V != out && ldo >= 2 * (int64_t)ncols, "pair_panel: bad arguments");
    if (nnodes * ncols == 0) return RVGP_OK;
    pair_panel_kernel<<<cdiv(nnodes * ncols, 256), 256, 0, h->stream>>>(nnodes, ncols, V, ldv, out, ldo);
    RVGP_LAUNCH_OK(h, "pair_panel_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_pair_combine_f64(rvgp_handle_t hh, int64_t nnodes, int ncols, const double* T, int64_t ldt, double alpha,
                                     double beta, double* W, int64_t ldw) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, nnodes >= 0 && ncols >= 0 && T != W && ldt >= 2 * (int64_t)ncols, "pair_combine: bad arguments");
    if (nnodes * ncols == 0) return RVGP_OK;
    pair_combine_kernel<<<cdiv(nnodes * ncols, 256), 256, 0, h->stream>>>(nnodes, ncols, T, ldt, alpha, beta, W, ldw);
    RVGP_LAUNCH_OK(h, "pair_combine_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_flip_odd_rows_f64(rvgp_handle_t hh, int64_t nnodes, int ncols, const int32_t* s, double* V, int64_t ldv) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, nnodes >= 0 && ncols >= 0 && s, "flip_odd_rows: bad arguments");
    if (nnodes * ncols == 0) return RVGP_OK;
    flip_odd_rows_kernel<<<cdiv(nnodes * ncols, 256), 256, 0, h->stream>>>(nnodes, ncols, s, V, ldv);
    RVGP_LAUNCH_OK(h, "flip_odd_rows_kernel");
    return RVGP_OK;
}

// K7 + K8: Procrustes connections and connection-Laplacian assembly; K11/a15 frame contractions.
//
// Replaces _parallel_transport_dijkstra (reference RVGP/lib/ptu_dijkstra.pyx:259-293): for every stored CSR
// entry (i,j) -- the diagonal included -- TtT = T_i^T T_j (d x d), dgesvd('A','A'), R = U * VT, i.e. the
// orthogonal polar factor of TtT (in O(d), det may be -1); and compute_connection_laplacian (reference
// RVGP/geometry.py:35-42): Lc block (i,i) = deg_i * R_ii, (i,j) = -R_ij with deg_i = #non-self neighbours.
//
// One thread per stored block; the d x d SVD is a one-sided Jacobi kept in registers (template on d).
// Gather-bound: 2*8*nnzb*D*d bytes of gauges read (mostly L2 hits), 8*nnzb*d^2 written.
#include "common.cuh"

namespace rvgp {

template <int d>
__device__ __forceinline__ void polar_factor(double (&B)[d][d], double (&R)[d][d]) {
    // one-sided Jacobi: B <- B V (columns orthogonal), V accumulated; U = normalised columns; R = U V^T
    double V[d][d];
#pragma unroll
    for (int p = 0; p < d; ++p)
#pragma unroll
        for (int q = 0; q < d; ++q) V[p][q] = (p == q) ? 1.0 : 0.0;
    if (d > 1) {
        for (int sweep = 0; sweep < 40; ++sweep) {
            bool rotated = false;
#pragma unroll
            for (int p = 0; p < d - 1; ++p)
#pragma unroll
                for (int q = p + 1; q < d; ++q) {
                    double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
                    for (int r = 0; r < d; ++r) { a = fma(B[r][p], B[r][p], a); b = fma(B[r][q], B[r][q], b); g = fma(B[r][p], B[r][q], g); }
                    if (fabs(g) > 1e-15 * sqrt(a * b) && g != 0.0) {
                        rotated = true;
                        const double zeta = (b - a) / (2.0 * g);
                        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
                        for (int r = 0; r < d; ++r) {
                            const double bp = B[r][p], bq = B[r][q];
                            B[r][p] = c * bp - s * bq; B[r][q] = s * bp + c * bq;
                            const double vp = V[r][p], vq = V[r][q];
                            V[r][p] = c * vp - s * vq; V[r][q] = s * vp + c * vq;
                        }
                    }
                }
            if (!rotated) break;
        }
    }
    // normalise columns of B -> U.  A (numerically) zero singular value leaves the direction undetermined
    // (LAPACK returns an arbitrary completion there too); complete by Gram-Schmidt against the other columns.
#pragma unroll
    for (int p = 0; p < d; ++p) {
        double nrm = 0.0;
#pragma unroll
        for (int r = 0; r < d; ++r) nrm = fma(B[r][p], B[r][p], nrm);
        nrm = sqrt(nrm);
        if (nrm > 1e-150) {
#pragma unroll
            for (int r = 0; r < d; ++r) B[r][p] /= nrm;
        } else {
            for (int e = 0; e < d; ++e) {
                double w[d];
#pragma unroll
                for (int r = 0; r < d; ++r) w[r] = (r == e) ? 1.0 : 0.0;
#pragma unroll
                for (int q = 0; q < d; ++q) {
                    if (q == p) continue;
                    double dot = 0.0;
#pragma unroll
                    for (int r = 0; r < d; ++r) dot = fma(B[r][q], w[r], dot);
#pragma unroll
                    for (int r = 0; r < d; ++r) w[r] -= dot * B[r][q];
                }
                double wn = 0.0;
#pragma unroll
                for (int r = 0; r < d; ++r) wn = fma(w[r], w[r], wn);
                if (wn > 0.25 / d) {
                    wn = sqrt(wn);
#pragma unroll
                    for (int r = 0; r < d; ++r) B[r][p] = w[r] / wn;
                    break;
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < d; ++p)
#pragma unroll
        for (int q = 0; q < d; ++q) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < d; ++r) s = fma(B[p][r], V[q][r], s);
            R[p][q] = s;
        }
}

template <int d>
__global__ void __launch_bounds__(128)
connections_kernel(const double* __restrict__ gauges, int n, int D, const int* __restrict__ indptr,
                   const int* __restrict__ indices, int64_t nnzb, double* __restrict__ Lc_vals, double* __restrict__ R_vals) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnzb) return;
    // row of entry e: largest i with indptr[i] <= e
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(indptr + mid) <= e) lo = mid; else hi = mid;
    }
    const int i = lo, j = __ldg(indices + e);
    const double* Ti = gauges + (int64_t)i * D * d;
    const double* Tj = gauges + (int64_t)j * D * d;
    double B[d][d], R[d][d];
#pragma unroll
    for (int p = 0; p < d; ++p)
#pragma unroll
        for (int q = 0; q < d; ++q) B[p][q] = 0.0;
    for (int k = 0; k < D; ++k) {          // pyx:262-267, sequential over the ambient coordinate
        double ti[d], tj[d];
#pragma unroll
        for (int p = 0; p < d; ++p) { ti[p] = __ldg(Ti + k * d + p); tj[p] = __ldg(Tj + k * d + p); }
#pragma unroll
        for (int p = 0; p < d; ++p)
#pragma unroll
            for (int q = 0; q < d; ++q) B[p][q] = fma(ti[p], tj[q], B[p][q]);
    }
    polar_factor<d>(B, R);
    const double scale = (i == j) ? (double)(__ldg(indptr + i + 1) - __ldg(indptr + i) - 1) : -1.0;
#pragma unroll
    for (int p = 0; p < d; ++p)
#pragma unroll
        for (int q = 0; q < d; ++q) {
            if (R_vals) R_vals[e * d * d + p * d + q] = R[p][q];
            if (Lc_vals) Lc_vals[e * d * d + p * d + q] = scale * R[p][q];
        }
}

// mode 0: out[i][q][c] = sum_p G[i][p][q] x[i][p][c]      (n,D,nc) -> (n,d,nc)
// mode 1: out[i][p][c] = sum_q G[i][p][q] x[i][q][c]      (n,d,nc) -> (n,D,nc)
__global__ void frame_apply_kernel(const double* __restrict__ G, int64_t n, int D, int d, const double* __restrict__ x,
                                   double* __restrict__ out, int nc, int mode, double scale) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int od = (mode == 0) ? d : D, id = (mode == 0) ? D : d;
    if (idx >= n * od * nc) return;
    const int c = (int)(idx % nc);
    const int o = (int)((idx / nc) % od);
    const int64_t i = idx / ((int64_t)nc * od);
    const double* g = G + i * D * d;
    double s = 0.0;
    for (int k = 0; k < id; ++k) {
        const double gv = (mode == 0) ? __ldg(g + k * d + o) : __ldg(g + o * d + k);
        s = fma(gv, __ldg(x + (i * id + k) * nc + c), s);
    }
    out[idx] = scale * s;
}

// Initial block for the Lc eigensolver from the scalar-Laplacian eigenvectors: smooth vector fields ~ (smooth scalar
// function) x (constant ambient direction projected on the tangent frame):
//   out[(i*d+q), c] = gauges[i][a][q] * U[i][j],   j = c / D, a = c % D
__global__ void lift_guess_kernel(const double* __restrict__ G, int64_t n, int D, int d, const double* __restrict__ U, int64_t ldu,
                                  int kL, double* __restrict__ out, int64_t ldo, int ncols) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * d * ncols) return;
    const int c = (int)(idx % ncols);
    const int64_t row = idx / ncols;
    const int64_t i = row / d;
    const int q = (int)(row % d);
    const int j = c / D, a = c % D;
    out[row * ldo + c] = (j < kL) ? __ldg(G + (i * D + a) * d + q) * __ldg(U + i * ldu + j) : 0.0;
}

}  // namespace rvgp

using namespace rvgp;

extern "C" int rvgp_lift_guess(rvgp_handle_t hh, const double* gauges, int64_t n, int D, int d, const double* U, int64_t ldu,
                               int kL, double* out, int64_t ldo, int ncols) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, ncols >= 0 && ncols <= kL * D, "lift_guess: ncols must be <= kL * D");
    const int64_t tot = n * d * ncols;
    if (tot == 0) return RVGP_OK;
    lift_guess_kernel<<<cdiv(tot, 256), 256, 0, h->stream>>>(gauges, n, D, d, U, ldu, kL, out, ldo, ncols);
    RVGP_LAUNCH_OK(h, "lift_guess_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_connections(rvgp_handle_t hh, const double* gauges, int n, int D, int d, const int32_t* indptr,
                                const int32_t* indices, int64_t nnzb, double* Lc_vals, double* R_vals) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, D >= d, "Embedding dimension must be less or equal to the ambient dimension of input data");
    RVGP_REQUIRE(h, d >= 1 && d <= 8, "connections: manifold dimension d must be in [1,8]");
    if (nnzb == 0) return RVGP_OK;
    const int grid = cdiv(nnzb, 128);
    switch (d) {
#define RVGP_CASE(DD) case DD: connections_kernel<DD><<<grid, 128, 0, h->stream>>>(gauges, n, D, indptr, indices, nnzb, Lc_vals, R_vals); break;
        RVGP_CASE(1) RVGP_CASE(2) RVGP_CASE(3) RVGP_CASE(4) RVGP_CASE(5) RVGP_CASE(6) RVGP_CASE(7) RVGP_CASE(8)
#undef RVGP_CASE
    }
    RVGP_LAUNCH_OK(h, "connections_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_frame_apply(rvgp_handle_t hh, const double* gauges, int64_t n, int D, int d, const double* x, double* out,
                                int ncols, int mode, double scale) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, mode == 0 || mode == 1, "frame_apply: mode must be 0 or 1");
    const int64_t tot = n * (mode == 0 ? d : D) * ncols;
    if (tot == 0) return RVGP_OK;
    frame_apply_kernel<<<cdiv(tot, 256), 256, 0, h->stream>>>(gauges, n, D, d, x, out, ncols, mode, scale);
    RVGP_LAUNCH_OK(h, "frame_apply_kernel");
    return RVGP_OK;
}

// K9: block-CSR SpMM with the Chebyshev three-term update fused into the epilogue.
//
//   Y = alpha * (A @ X) + beta * X + gamma * W
//
// Replaces the ARPACK mat-vec inside scipy.sparse.linalg.eigsh (reference RVGP/geometry.py:73).
// HBM-bound: per launch the algorithmic traffic is
//     nnzb*(8 d^2 + 4) + 4 (nbrows+1)            matrix stream (read once)
//   + 8 * nbrows*d*ncols * (2 + [gamma != 0])    X read once, Y written once, W read once
// (DESIGN.md section "K9").  Layout: a group of LPR lanes owns one block row and walks its stored
// blocks; lanes run across the columns of the block vector so every gather of a neighbour's rows is a
// contiguous LPR*8-byte segment.  Nodes are expected in a locality-preserving order (Morton) so the
// neighbour gathers of the 64 consecutive block rows of a CTA hit L1/L2.
#include "common.cuh"

namespace rvgp {

template <int D, int LPR, int CPL, int U, bool PATTERN>
__global__ void __launch_bounds__(256)
bsr_spmm_kernel(int nbrows, const int* __restrict__ indptr, const int* __restrict__ indices,
                const double* __restrict__ vals, const double* __restrict__ X, int64_t ldx,
                const double* __restrict__ W, int64_t ldw, double* __restrict__ Y, int64_t ldy,
                int ncols, double alpha, double beta, double gamma, int rows_per_group) {
    constexpr int GPW = 32 / LPR;          // row groups per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LPR, l = lane % LPR;
    const int rows_per_cta = 8 * GPW * rows_per_group;
    const int row0 = blockIdx.x * rows_per_cta;

    bool colok[CPL];
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) colok[cc] = (l + LPR * cc) < ncols;

    for (int it = 0; it < rows_per_group; ++it) {
        // interleave the warps of a CTA over consecutive rows: concurrent warps touch neighbouring nodes
        const int i = row0 + it * (8 * GPW) + warp * GPW + g;
        if (i >= nbrows) continue;
        const int e0 = __ldg(indptr + i), e1 = __ldg(indptr + i + 1);
        double acc[D][CPL];
#pragma unroll
        for (int p = 0; p < D; ++p)
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) acc[p][cc] = 0.0;

        for (int e = e0; e < e1; e += U) {
            int j[U];
#pragma unroll
            for (int u = 0; u < U; ++u) j[u] = (e + u < e1) ? __ldg(indices + e + u) : -1;
            double x[U][D][CPL];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int q = 0; q < D; ++q)
#pragma unroll
                    for (int cc = 0; cc < CPL; ++cc)
                        x[u][q][cc] = (j[u] >= 0 && colok[cc])
                                          ? __ldg(X + ((int64_t)j[u] * D + q) * ldx + l + LPR * cc)
                                          : 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (j[u] < 0) continue;
                if (PATTERN) {
                    // scalar graph Laplacian D - A with unit weights (geometry.py:61):
                    // diagonal entry = number of non-self neighbours, every other stored entry = -1
                    const double r = (j[u] == i) ? (double)(e1 - e0 - 1) : -1.0;
#pragma unroll
                    for (int cc = 0; cc < CPL; ++cc) acc[0][cc] = fma(r, x[u][0][cc], acc[0][cc]);
                } else {
                    const double* rp = vals + (int64_t)(e + u) * (D * D);
#pragma unroll
                    for (int p = 0; p < D; ++p)
#pragma unroll
                        for (int q = 0; q < D; ++q) {
                            const double r = __ldg(rp + p * D + q);
#pragma unroll
                            for (int cc = 0; cc < CPL; ++cc) acc[p][cc] = fma(r, x[u][q][cc], acc[p][cc]);
                        }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < D; ++p)
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                if (!colok[cc]) continue;
                const int c = l + LPR * cc;
                const int64_t r = (int64_t)i * D + p;
                double y = alpha * acc[p][cc];
                if (beta != 0.0) y = fma(beta, __ldg(X + r * ldx + c), y);
                if (gamma != 0.0) y = fma(gamma, __ldg(W + r * ldw + c), y);
                Y[r * ldy + c] = y;
            }
    }
}

template <int D, bool PATTERN>
static int launch_spmm_d(Handle* h, int nbrows, const int* indptr, const int* indices, const double* vals,
                         const double* X, int64_t ldx, const double* W, int64_t ldw, double* Y, int64_t ldy,
                         int ncols, double alpha, double beta, double gamma) {
    constexpr int U = (D <= 2) ? 4 : (D <= 3 ? 2 : 1);
    const int rpg = 8;  // rows per lane-group per CTA -> 64 * (32/LPR) consecutive block rows per CTA
#define RVGP_SPMM_LAUNCH(LPR, CPL)                                                                     \
    do {                                                                                               \
        const int rows_per_cta = 8 * (32 / LPR) * rpg;                                                 \
        const int grid = cdiv(nbrows, rows_per_cta);                                                   \
        bsr_spmm_kernel<D, LPR, CPL, U, PATTERN><<<grid, 256, 0, h->stream>>>(                         \
            nbrows, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma, rpg);    \
    } while (0)
    if (ncols <= 8) RVGP_SPMM_LAUNCH(8, 1);
    else if (ncols <= 16) RVGP_SPMM_LAUNCH(16, 1);
    else if (ncols <= 32) RVGP_SPMM_LAUNCH(32, 1);
    else RVGP_SPMM_LAUNCH(32, 2);
#undef RVGP_SPMM_LAUNCH
    RVGP_LAUNCH_OK(h, "bsr_spmm_kernel");
    return RVGP_OK;
}

static int spmm_dispatch(Handle* h, int nbrows, int d, const int* indptr, const int* indices, const double* vals,
                         const double* X, int64_t ldx, const double* W, int64_t ldw, double* Y, int64_t ldy,
                         int ncols, double alpha, double beta, double gamma) {
    RVGP_REQUIRE(h, nbrows >= 0 && ncols >= 1 && ncols <= 64, "spmm: ncols must be in [1,64]");
    RVGP_REQUIRE(h, gamma == 0.0 || W != nullptr, "spmm: W required when gamma != 0");
    RVGP_REQUIRE(h, Y != X && Y != W, "spmm: Y must not alias X or W");
    if (nbrows == 0) return RVGP_OK;
    if (vals == nullptr) {
        RVGP_REQUIRE(h, d == 1, "spmm: pattern mode (vals == NULL) needs d == 1");
        return launch_spmm_d<1, true>(h, nbrows, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma);
    }
    switch (d) {
#define RVGP_CASE(DD) case DD: return launch_spmm_d<DD, false>(h, nbrows, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma);
        RVGP_CASE(1) RVGP_CASE(2) RVGP_CASE(3) RVGP_CASE(4) RVGP_CASE(5) RVGP_CASE(6) RVGP_CASE(7) RVGP_CASE(8)
#undef RVGP_CASE
        default: return set_error(h, RVGP_ERR_BAD_ARG, "spmm: block size d must be in [1,8]%s%s");
    }
}

}  // namespace rvgp

using namespace rvgp;

extern "C" int rvgp_bsr_spmm_f64(rvgp_handle_t hh, int nbrows, int d, const int32_t* indptr, const int32_t* indices,
                                 const double* vals, const double* X, int64_t ldx, const double* W, int64_t ldw,
                                 double* Y, int64_t ldy, int ncols, double alpha, double beta, double gamma) {
    Handle* h = H(hh);
    return spmm_dispatch(h, nbrows, d, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma);
}

// Scaled Chebyshev filter (Zhou & Saad, "A Chebyshev-Davidson algorithm", Alg. 3.1 form):
//   e = (hi - lo_cut)/2, c = (hi + lo_cut)/2, sigma_1 = e / (lo_spec - c), tau = 2 / sigma_1
//   Y_1     = (A V - c V) * sigma_1 / e
//   Y_{i+1} = 2 sigma_{i+1}/e (A Y_i - c Y_i) - sigma_i sigma_{i+1} Y_{i-1},   sigma_{i+1} = 1/(tau - sigma_i)
// Every step is ONE fused SpMM launch.  The three block vectors rotate; the result is copied back into V
// when it does not land there.
extern "C" int rvgp_cheb_filter_f64(rvgp_handle_t hh, int nbrows, int d, const int32_t* indptr, const int32_t* indices,
                                    const double* vals, double* V, int64_t ldv, double* work0, double* work1,
                                    int64_t ldw, int ncols, int degree, double lo_spec, double lo_cut, double hi) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, degree >= 0, "cheb_filter: degree >= 0");
    RVGP_REQUIRE(h, hi > lo_cut && lo_cut > lo_spec, "cheb_filter: need lo_spec < lo_cut < hi");
    if (degree == 0 || nbrows == 0) return RVGP_OK;
    const double e = 0.5 * (hi - lo_cut), c = 0.5 * (hi + lo_cut);
    const double sigma1 = e / (lo_spec - c), tau = 2.0 / sigma1;
    double sigma = sigma1;
    double* buf[3] = {V, work0, work1};
    int64_t ld[3] = {ldv, ldw, ldw};
    int rc = spmm_dispatch(h, nbrows, d, indptr, indices, vals, buf[0], ld[0], nullptr, 0, buf[1], ld[1], ncols,
                           sigma1 / e, -c * sigma1 / e, 0.0);
    if (rc) return rc;
    int prev = 0, cur = 1;
    for (int i = 2; i <= degree; ++i) {
        const double sn = 1.0 / (tau - sigma);
        const int nxt = 3 - prev - cur;
        rc = spmm_dispatch(h, nbrows, d, indptr, indices, vals, buf[cur], ld[cur], buf[prev], ld[prev], buf[nxt],
                           ld[nxt], ncols, 2.0 * sn / e, -2.0 * sn * c / e, -sigma * sn);
        if (rc) return rc;
        sigma = sn;
        prev = cur;
        cur = nxt;
    }
    if (cur != 0) {
        RVGP_CUDA_OK(h, cudaMemcpy2DAsync(V, ldv * sizeof(double), buf[cur], ld[cur] * sizeof(double),
                                          (size_t)ncols * sizeof(double), (size_t)nbrows * d,
                                          cudaMemcpyDeviceToDevice, h->stream));
    }
    return RVGP_OK;
}

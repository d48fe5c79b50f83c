// K9 (v3): row-group merged block-SpMM.
//
// ncu on v2 (profiles/r01_spmm_v2_*): DRAM traffic == algorithmic bytes, but the L1 data pipe is ~76 % busy and 8 of
// the ~13 wavefronts per stored block are the gather of the neighbour's rows of X.  Adjacent (Morton-ordered) block
// rows share most of their neighbours, so this version processes R consecutive block rows TOGETHER:
//   * a one-off plan merges the R sorted column lists of every row group into one union list; each union entry is
//     (column, R-bit mask of the rows that really store that column);
//   * the kernel walks the union list: ONE gather of the neighbour's X rows (LDG.128 per lane) feeds all rows whose
//     mask bit is set; every row keeps a running pointer into its own CSR entries, so block values are read exactly
//     once and nothing is duplicated or zero-padded.
// With R = 4 the union holds ~half as many entries as the four rows together, i.e. half the X gathers through L1.
// Y = alpha * (A @ X) + beta * X + gamma * W; needs even ncols / leading dimensions, 16-byte aligned buffers.
#include <cub/cub.cuh>

#include "common.cuh"

namespace rvgp {

// ---- plan -------------------------------------------------------------------------------------------------------
template <int R>
__global__ void merge_count_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, int n, int ngroups,
                                   int* __restrict__ ulen, int2* __restrict__ uent, const int* __restrict__ gptr) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    int p[R], pe[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = g * R + r;
        p[r] = (row < n) ? indptr[row] : 0;
        pe[r] = (row < n) ? indptr[row + 1] : 0;
    }
    int cnt = 0;
    const int base = gptr ? gptr[g] : 0;
    for (;;) {
        int mn = 0x7fffffff;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (p[r] < pe[r]) mn = min(mn, indices[p[r]] & 0x7fffffff);
        if (mn == 0x7fffffff) break;
        int mask = 0;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (p[r] < pe[r] && (indices[p[r]] & 0x7fffffff) == mn) { mask |= 1 << r; ++p[r]; }
        if (uent) uent[base + cnt] = make_int2(mn, mask);
        ++cnt;
    }
    if (ulen) ulen[g] = cnt;
}

// ---- SpMM ---------------------------------------------------------------------------------------------------------
// One group of LPR lanes per row group; lane l owns column pairs l + LPR*cc (double2).
template <int D, int R, int LPR, int CPL2, bool PATTERN, bool ROT2>
__global__ void __launch_bounds__(256, (R * D * CPL2 <= 8) ? 4 : 2)
bsr_spmm_merged_kernel(int nbrows, int ngroups, const int* __restrict__ indptr, const int* __restrict__ indices,
                       const int* __restrict__ gptr, const int2* __restrict__ uent, const double* __restrict__ vals,
                       const double* __restrict__ X, int64_t ldx, const double* __restrict__ W, int64_t ldw,
                       double* __restrict__ Y, int64_t ldy, int ncols, double alpha, double beta, double gamma,
                       int groups_per_lanegroup) {
    constexpr int GPW = 32 / LPR;
    constexpr int VB = ROT2 ? 2 : D * D;           // doubles per stored block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPR, l = lane % LPR;
    const int npairs = ncols >> 1;
    bool colok[CPL2];
#pragma unroll
    for (int cc = 0; cc < CPL2; ++cc) colok[cc] = (l + LPR * cc) < npairs;
    const int per_cta = 8 * GPW * groups_per_lanegroup;
    const int g0 = blockIdx.x * per_cta;

    for (int it = 0; it < groups_per_lanegroup; ++it) {
        const int g = g0 + it * (8 * GPW) + warp * GPW + sub;
        if (g >= ngroups) continue;
        int p[R];                                   // running CSR entry pointer of every row of the group
        int deg[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = g * R + r;
            p[r] = (row < nbrows) ? __ldg(indptr + row) : 0;
            deg[r] = (row < nbrows) ? (__ldg(indptr + row + 1) - p[r] - 1) : 0;
        }
        double2 acc[R][D][CPL2];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int q = 0; q < D; ++q)
#pragma unroll
                for (int cc = 0; cc < CPL2; ++cc) acc[r][q][cc] = make_double2(0.0, 0.0);

        const int k0 = __ldg(gptr + g), k1 = __ldg(gptr + g + 1);
        for (int k = k0; k < k1; k += 2) {
            int2 en[2];
            en[0] = __ldg(uent + k);
            en[1] = (k + 1 < k1) ? __ldg(uent + k + 1) : make_int2(0, 0);
            double2 x[2][D][CPL2];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int q = 0; q < D; ++q)
#pragma unroll
                    for (int cc = 0; cc < CPL2; ++cc)
                        x[u][q][cc] = (en[u].y != 0 && colok[cc])
                                          ? __ldg(reinterpret_cast<const double2*>(X + ((int64_t)en[u].x * D + q) * ldx) + l + LPR * cc)
                                          : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int mask = en[u].y;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (!((mask >> r) & 1)) continue;          // warp-group uniform
                    if (PATTERN) {
                        const double rr = (en[u].x == g * R + r) ? (double)deg[r] : -1.0;
#pragma unroll
                        for (int cc = 0; cc < CPL2; ++cc) {
                            acc[r][0][cc].x = fma(rr, x[u][0][cc].x, acc[r][0][cc].x);
                            acc[r][0][cc].y = fma(rr, x[u][0][cc].y, acc[r][0][cc].y);
                        }
                    } else {
                        const double* rp = vals + (int64_t)p[r] * VB;
                        double m[D * D];
                        if (ROT2) {
                            const double2 ab = __ldg(reinterpret_cast<const double2*>(rp));
                            const double sg = (__ldg(indices + p[r]) < 0) ? -1.0 : 1.0;
                            m[0] = ab.x; m[1 % (D * D)] = -sg * ab.y; m[2 % (D * D)] = ab.y; m[3 % (D * D)] = sg * ab.x;
                        } else if ((D * D) % 2 == 0) {
#pragma unroll
                            for (int v = 0; v < D * D; v += 2) {
                                const double2 rv = __ldg(reinterpret_cast<const double2*>(rp + v));
                                m[v] = rv.x; m[v + 1] = rv.y;
                            }
                        } else {
#pragma unroll
                            for (int v = 0; v < D * D; ++v) m[v] = __ldg(rp + v);
                        }
#pragma unroll
                        for (int pp = 0; pp < D; ++pp)
#pragma unroll
                            for (int q = 0; q < D; ++q)
#pragma unroll
                                for (int cc = 0; cc < CPL2; ++cc) {
                                    acc[r][pp][cc].x = fma(m[pp * D + q], x[u][q][cc].x, acc[r][pp][cc].x);
                                    acc[r][pp][cc].y = fma(m[pp * D + q], x[u][q][cc].y, acc[r][pp][cc].y);
                                }
                    }
                    ++p[r];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = g * R + r;
            if (row >= nbrows) continue;
#pragma unroll
            for (int pp = 0; pp < D; ++pp)
#pragma unroll
                for (int cc = 0; cc < CPL2; ++cc) {
                    if (!colok[cc]) continue;
                    const int c2 = l + LPR * cc;
                    const int64_t rr = (int64_t)row * D + pp;
                    double2 y = make_double2(alpha * acc[r][pp][cc].x, alpha * acc[r][pp][cc].y);
                    if (beta != 0.0) {
                        const double2 xv = __ldg(reinterpret_cast<const double2*>(X + rr * ldx) + c2);
                        y.x = fma(beta, xv.x, y.x); y.y = fma(beta, xv.y, y.y);
                    }
                    if (gamma != 0.0) {
                        const double2 wv = __ldg(reinterpret_cast<const double2*>(W + rr * ldw) + c2);
                        y.x = fma(gamma, wv.x, y.x); y.y = fma(gamma, wv.y, y.y);
                    }
                    reinterpret_cast<double2*>(Y + rr * ldy)[c2] = y;
                }
        }
    }
}

template <int D, int R, bool PATTERN, bool ROT2>
static int launch_merged(Handle* h, int nbrows, const int* indptr, const int* indices, const int* gptr, const int2* uent,
                         const double* vals, const double* X, int64_t ldx, const double* W, int64_t ldw, double* Y,
                         int64_t ldy, int ncols, double alpha, double beta, double gamma) {
    const int ngroups = cdiv(nbrows, R);
    const int npairs = ncols >> 1;
    const int gpl = 4;
#define RVGP_M(LPR, CPL2)                                                                                             \
    do {                                                                                                              \
        const int per_cta = 8 * (32 / LPR) * gpl;                                                                     \
        bsr_spmm_merged_kernel<D, R, LPR, CPL2, PATTERN, ROT2><<<cdiv(ngroups, per_cta), 256, 0, h->stream>>>(       \
            nbrows, ngroups, indptr, indices, gptr, uent, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma, gpl); \
    } while (0)
    int lpr = h->spmm_lpr;
    if (lpr != 8 && lpr != 16 && lpr != 32) lpr = (npairs <= 8) ? 8 : (npairs <= 16 ? 16 : 32);
    while (lpr < 32 && lpr * 2 < npairs) lpr *= 2;
    const int cpl = (npairs + lpr - 1) / lpr;
    if (lpr == 8) { if (cpl <= 1) RVGP_M(8, 1); else RVGP_M(8, 2); }
    else if (lpr == 16) { if (cpl <= 1) RVGP_M(16, 1); else RVGP_M(16, 2); }
    else RVGP_M(32, 1);
#undef RVGP_M
    RVGP_LAUNCH_OK(h, "bsr_spmm_merged_kernel");
    return RVGP_OK;
}

int spmm_merged_dispatch(Handle* h, int nbrows, int d, int R, const int* indptr, const int* indices, const int* gptr,
                         const int* uent, const double* vals, const double* X, int64_t ldx, const double* W, int64_t ldw,
                         double* Y, int64_t ldy, int ncols, double alpha, double beta, double gamma) {
    RVGP_REQUIRE(h, ncols >= 2 && ncols <= 64 && ncols % 2 == 0, "spmm_merged: ncols must be even and in [2,64]");
    RVGP_REQUIRE(h, ldx % 2 == 0 && ldy % 2 == 0 && (W == nullptr || ldw % 2 == 0) && (uintptr_t)X % 16 == 0 &&
                        (uintptr_t)Y % 16 == 0 && (uintptr_t)W % 16 == 0 && (uintptr_t)vals % 16 == 0,
                 "spmm_merged: buffers must be 16-byte aligned with even leading dimensions");
    RVGP_REQUIRE(h, gamma == 0.0 || W != nullptr, "spmm_merged: W required when gamma != 0");
    RVGP_REQUIRE(h, Y != X && Y != W, "spmm_merged: Y must not alias X or W");
    RVGP_REQUIRE(h, R == 4 || R == 8, "spmm_merged: R must be 4 or 8");
    if (nbrows == 0) return RVGP_OK;
    const int2* ue = reinterpret_cast<const int2*>(uent);
#define RVGP_GO(DD, PAT, ROT)                                                                                          \
    return (R == 4) ? launch_merged<DD, 4, PAT, ROT>(h, nbrows, indptr, indices, gptr, ue, vals, X, ldx, W, ldw, Y, ldy, \
                                                     ncols, alpha, beta, gamma)                                        \
                    : launch_merged<DD, 8, PAT, ROT>(h, nbrows, indptr, indices, gptr, ue, vals, X, ldx, W, ldw, Y, ldy, \
                                                     ncols, alpha, beta, gamma)
    if (vals == nullptr) { RVGP_REQUIRE(h, d == 1, "spmm_merged: pattern mode needs d == 1"); RVGP_GO(1, true, false); }
    if (d == -2) RVGP_GO(2, false, true);
    if (d == 1) RVGP_GO(1, false, false);
    if (d == 2) RVGP_GO(2, false, false);
    if (d == 3) RVGP_GO(3, false, false);
#undef RVGP_GO
    return set_error(h, RVGP_ERR_BAD_ARG, "spmm_merged: block size d must be 1, 2, 3 or -2 (rot2)%s%s");
}

}  // namespace rvgp

using namespace rvgp;

extern "C" int64_t rvgp_bsr_merge_plan_workspace_bytes(int nbrows, int R) {
    const int ngroups = (nbrows + R - 1) / R;
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (int*)nullptr, (int*)nullptr, ngroups + 1);
    return (int64_t)(((size_t)(ngroups + 1) * 4 + 255) / 256 * 256 + (b + 255) / 256 * 256);
}

// Row-group merge plan.  Pass 1 (uent == NULL): fills gptr (ngroups+1, exclusive prefix of the union lengths); read
// gptr[ngroups] for the total number of union entries, allocate uent (int32 pairs (column, row mask)), then call again
// with uent to fill it.  `indices` may carry the ROT2 flip bit (ignored here).
extern "C" int rvgp_bsr_merge_plan(rvgp_handle_t hh, int nbrows, const int32_t* indptr, const int32_t* indices, int R,
                                   int32_t* gptr, int32_t* uent, void* workspace, int64_t workspace_bytes) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, R == 4 || R == 8, "merge_plan: R must be 4 or 8");
    const int ngroups = cdiv(nbrows, R);
    if (ngroups == 0) return RVGP_OK;
    if (uent == nullptr) {
        if (rvgp_bsr_merge_plan_workspace_bytes(nbrows, R) > workspace_bytes)
            return set_error(h, RVGP_ERR_CAPACITY, "merge_plan: workspace too small%s%s");
        int* ulen = (int*)workspace;
        void* cubtmp = (char*)workspace + ((size_t)(ngroups + 1) * 4 + 255) / 256 * 256;
        size_t cb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, cb, (int*)nullptr, (int*)nullptr, ngroups + 1);
        RVGP_CUDA_OK(h, cudaMemsetAsync(ulen + ngroups, 0, sizeof(int), h->stream));
        if (R == 4) merge_count_kernel<4><<<cdiv(ngroups, 128), 128, 0, h->stream>>>(indptr, indices, nbrows, ngroups, ulen, nullptr, nullptr);
        else merge_count_kernel<8><<<cdiv(ngroups, 128), 128, 0, h->stream>>>(indptr, indices, nbrows, ngroups, ulen, nullptr, nullptr);
        RVGP_LAUNCH_OK(h, "merge_count_kernel");
        RVGP_CUDA_OK(h, cub::DeviceScan::ExclusiveSum(cubtmp, cb, ulen, gptr, ngroups + 1, h->stream));
        h->launches++;
    } else {
        if (R == 4) merge_count_kernel<4><<<cdiv(ngroups, 128), 128, 0, h->stream>>>(indptr, indices, nbrows, ngroups, nullptr, reinterpret_cast<int2*>(uent), gptr);
        else merge_count_kernel<8><<<cdiv(ngroups, 128), 128, 0, h->stream>>>(indptr, indices, nbrows, ngroups, nullptr, reinterpret_cast<int2*>(uent), gptr);
        RVGP_LAUNCH_OK(h, "merge_count_kernel");
    }
    return RVGP_OK;
}

extern "C" int rvgp_bsr_spmm_merged_f64(rvgp_handle_t hh, int nbrows, int d, int R, const int32_t* indptr, const int32_t* indices,
                                        const int32_t* gptr, const int32_t* uent, const double* vals, const double* X,
                                        int64_t ldx, const double* W, int64_t ldw, double* Y, int64_t ldy, int ncols,
                                        double alpha, double beta, double gamma) {
    return spmm_merged_dispatch(H(hh), nbrows, d, R, indptr, indices, gptr, uent, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha,
                                beta, gamma);
}

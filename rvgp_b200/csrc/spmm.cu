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

#ifndef RVGP_V2_NT
#define RVGP_V2_NT 256
#endif
#ifndef RVGP_SPMM_STREAMING
#define RVGP_SPMM_STREAMING 1
#endif
#if RVGP_SPMM_STREAMING
#define RVGP_STREAM_LOAD2(p) __ldcs(p)
#define RVGP_STREAM_STORE2(p, v) __stcs(p, v)
#else
#define RVGP_STREAM_LOAD2(p) __ldg(p)
#define RVGP_STREAM_STORE2(p, v) (*(p) = (v))
#endif

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

// ---- v2: 128-bit loads ---------------------------------------------------------------------------------------
// ncu (profiles/r01_spmm_*): the L1 data pipe moves 64 B per wavefront for 64-bit-per-lane loads, so the v1 kernel is
// capped near 64 B/clk/SM.  Here every lane owns PAIRS of adjacent columns and gathers them with LDG.128 (double2),
// and even-sized value blocks are read as double2 as well: twice the bytes per wavefront, half the load instructions.
// Needs even ncols / leading dimensions and 16-byte aligned X, W, Y (the dispatcher falls back to v1 otherwise).
//
// ROT2 (d == 2 only): every block of the connection Laplacian is a scaled 2x2 orthogonal matrix
//   [[a, -s*b], [b, s*a]],  s = +-1  (rotation or reflection; the diagonal block is deg * I),
// so it can be stored as ONE double2 (a, b) with s in the sign bit of the column index: half the matrix bytes and one
// LDG.128 per stored block instead of two.  Roofline accounting keeps the uncompressed formula.
//
// STAGE: a warp-uniform LDG.128 still costs four L1 data-pipe wavefronts (quarter-warp granularity), so the per-block value
// loads were ~40 % of the pipe work.  With STAGE the CTA first copies the (contiguous) values and column indices of its 64
// block rows into shared memory with coalesced loads; the inner loop then reads them with broadcast LDS (one wavefront).
template <int D, int LPR, int CPL2, int U, bool PATTERN, bool ROT2, bool STAGE, bool ROWLIST>
__global__ void __launch_bounds__(RVGP_V2_NT, (RVGP_V2_NT == 256 && D == 2 && CPL2 == 1 && !PATTERN && !ROT2 && !STAGE && !ROWLIST) ? 6 : 1)   // 40 regs, 48 warps/SM: +5 % (measured)
bsr_spmm_v2_kernel(int nbrows, const int* __restrict__ indptr, const int* __restrict__ indices,
                   const double* __restrict__ vals, const double* __restrict__ X, int64_t ldx,
                   const double* __restrict__ W, int64_t ldw, double* __restrict__ Y, int64_t ldy,
                   int ncols, double alpha, double beta, double gamma, int rows_per_group, int stage_cap,
                   const int* __restrict__ rowlist, int nlist) {
    constexpr int GPW = 32 / LPR;
    constexpr int VB = ROT2 ? 2 : D * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LPR, l = lane % LPR;
    const int rows_per_cta = (RVGP_V2_NT / 32) * GPW * rows_per_group;
    const int row0 = blockIdx.x * rows_per_cta;
    const int npairs = ncols >> 1;
    extern __shared__ __align__(16) double stage_smem[];
    double* svals = stage_smem;
    int* sidx = reinterpret_cast<int*>(stage_smem + (size_t)stage_cap * VB);
    int te0 = 0;
    bool staged = false;
    if (STAGE) {
        te0 = __ldg(indptr + row0);
        const int rend = (row0 + rows_per_cta < nbrows) ? row0 + rows_per_cta : nbrows;
        const int ne = __ldg(indptr + rend) - te0;
        staged = ne <= stage_cap;
        if (staged) {
            const double* vsrc = vals + (int64_t)te0 * VB;
            for (int t = threadIdx.x; t < ne * VB; t += RVGP_V2_NT) svals[t] = __ldg(vsrc + t);
            for (int t = threadIdx.x; t < ne; t += RVGP_V2_NT) sidx[t] = __ldg(indices + te0 + t);
        }
        __syncthreads();
    }
    bool colok[CPL2];
#pragma unroll
    for (int cc = 0; cc < CPL2; ++cc) colok[cc] = (l + LPR * cc) < npairs;

    for (int it = 0; it < rows_per_group; ++it) {
        const int vi = row0 + it * ((RVGP_V2_NT / 32) * GPW) + warp * GPW + g;
        if (vi >= (ROWLIST ? nlist : nbrows)) continue;
        const int i = ROWLIST ? __ldg(rowlist + vi) : vi;      // optional row list (boundary rows of a row-sharded matrix)
        const int e0 = __ldg(indptr + i), e1 = __ldg(indptr + i + 1);
        double2 acc[D][CPL2];
#pragma unroll
        for (int p = 0; p < D; ++p)
#pragma unroll
            for (int cc = 0; cc < CPL2; ++cc) acc[p][cc] = make_double2(0.0, 0.0);

        for (int e = e0; e < e1; e += U) {
            int j[U];
            bool flip[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                j[u] = (e + u < e1) ? ((STAGE && staged) ? sidx[e + u - te0] : __ldg(indices + e + u)) : -1;
                flip[u] = false;
                if (ROT2 && e + u < e1) { flip[u] = j[u] < 0; j[u] &= 0x7fffffff; }
            }
            double2 x[U][D][CPL2];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int q = 0; q < D; ++q)
#pragma unroll
                    for (int cc = 0; cc < CPL2; ++cc)
                        x[u][q][cc] = (j[u] >= 0 && colok[cc])
                                          ? __ldg(reinterpret_cast<const double2*>(X + ((int64_t)j[u] * D + q) * ldx) + l + LPR * cc)
                                          : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (j[u] < 0) continue;
                if (PATTERN) {
                    const double r = (j[u] == i) ? (double)(e1 - e0 - 1) : -1.0;
#pragma unroll
                    for (int cc = 0; cc < CPL2; ++cc) {
                        acc[0][cc].x = fma(r, x[u][0][cc].x, acc[0][cc].x);
                        acc[0][cc].y = fma(r, x[u][0][cc].y, acc[0][cc].y);
                    }
                } else {
                    double r[D * D];
                    if (STAGE && staged) {
                        const double* rp = svals + (size_t)(e + u - te0) * VB;      // broadcast LDS
                        if (ROT2) {
                            const double2 ab = *reinterpret_cast<const double2*>(rp);
                            const double sg = flip[u] ? -1.0 : 1.0;
                            r[0] = ab.x; r[1 % (D * D)] = -sg * ab.y; r[2 % (D * D)] = ab.y; r[3 % (D * D)] = sg * ab.x;
                        } else if ((D * D) % 2 == 0) {
#pragma unroll
                            for (int v = 0; v < D * D; v += 2) {
                                const double2 rv = *reinterpret_cast<const double2*>(rp + v);
                                r[v] = rv.x; r[v + 1] = rv.y;
                            }
                        } else {
#pragma unroll
                            for (int v = 0; v < D * D; ++v) r[v] = rp[v];
                        }
                    } else {
                        const double* rp = vals + (int64_t)(e + u) * VB;
                        if (ROT2) {
                            const double2 ab = __ldg(reinterpret_cast<const double2*>(rp));
                            const double sg = flip[u] ? -1.0 : 1.0;
                            r[0] = ab.x; r[1 % (D * D)] = -sg * ab.y; r[2 % (D * D)] = ab.y; r[3 % (D * D)] = sg * ab.x;
                        } else if ((D * D) % 2 == 0) {
#pragma unroll
                            for (int v = 0; v < D * D; v += 2) {
                                const double2 rv = __ldg(reinterpret_cast<const double2*>(rp + v));
                                r[v] = rv.x; r[v + 1] = rv.y;
                            }
                        } else {
#pragma unroll
                            for (int v = 0; v < D * D; ++v) r[v] = __ldg(rp + v);
                        }
                    }
#pragma unroll
                    for (int p = 0; p < D; ++p)
#pragma unroll
                        for (int q = 0; q < D; ++q)
#pragma unroll
                            for (int cc = 0; cc < CPL2; ++cc) {
                                acc[p][cc].x = fma(r[p * D + q], x[u][q][cc].x, acc[p][cc].x);
                                acc[p][cc].y = fma(r[p * D + q], x[u][q][cc].y, acc[p][cc].y);
                            }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < D; ++p)
#pragma unroll
            for (int cc = 0; cc < CPL2; ++cc) {
                if (!colok[cc]) continue;
                const int c2 = l + LPR * cc;
                const int64_t r = (int64_t)i * D + p;
                double2 y = make_double2(alpha * acc[p][cc].x, alpha * acc[p][cc].y);
                if (beta != 0.0) {
                    const double2 xv = __ldg(reinterpret_cast<const double2*>(X + r * ldx) + c2);
                    y.x = fma(beta, xv.x, y.x); y.y = fma(beta, xv.y, y.y);
                }
                if (gamma != 0.0) {
                    const double2* wp = reinterpret_cast<const double2*>(W + r * ldw) + c2;
                    // W is read once: for the value-free scalar Laplacian a streaming load (evict-first) helps (+6 %), for d=2 it does not
                    const double2 wv = PATTERN ? RVGP_STREAM_LOAD2(wp) : __ldg(wp);
                    y.x = fma(gamma, wv.x, y.x); y.y = fma(gamma, wv.y, y.y);
                }
                if (PATTERN) RVGP_STREAM_STORE2(reinterpret_cast<double2*>(Y + r * ldy) + c2, y);
                else reinterpret_cast<double2*>(Y + r * ldy)[c2] = y;
            }
    }
}

template <int D, bool PATTERN, bool ROT2>
static int launch_spmm_v2(Handle* h, int nbrows, const int* indptr, const int* indices, const double* vals,
                          const double* X, int64_t ldx, const double* W, int64_t ldw, double* Y, int64_t ldy,
                          int ncols, double alpha, double beta, double gamma) {
    const int rpg = 8;
    const int npairs = ncols >> 1;
#define RVGP_V2(LPR, CPL2)                                                                                 \
    do {                                                                                                   \
        constexpr int U = (D * CPL2 <= 2) ? 4 : ((D * CPL2 <= 4) ? 2 : 1);                                 \
        const int rows_per_cta = (RVGP_V2_NT / 32) * (32 / LPR) * rpg;                                                     \
        const int grid_ = cdiv(h->rowlist ? h->nlist : nbrows, rows_per_cta);                              \
        if (h->rowlist) {                                                                                  \
            bsr_spmm_v2_kernel<D, LPR, CPL2, U, PATTERN, ROT2, false, true><<<grid_, RVGP_V2_NT, 0, h->stream>>>( \
                nbrows, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma, rpg, 0, h->rowlist, h->nlist); \
        } else if (!PATTERN && h->spmm_stage) {                                                            \
            constexpr int VB_ = ROT2 ? 2 : D * D;                                                          \
            const int cap = rows_per_cta * 16;                                                             \
            const int smem = cap * (VB_ * 8 + 4);                                                          \
            auto kern = bsr_spmm_v2_kernel<D, LPR, CPL2, U, PATTERN, ROT2, true, false>;                   \
            if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
            kern<<<grid_, RVGP_V2_NT, smem, h->stream>>>(                                                         \
                nbrows, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma, rpg, cap, nullptr, 0); \
        } else {                                                                                           \
            bsr_spmm_v2_kernel<D, LPR, CPL2, U, PATTERN, ROT2, false, false><<<grid_, RVGP_V2_NT, 0, h->stream>>>( \
                nbrows, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma, rpg, 0, nullptr, 0); \
        }                                                                                                  \
    } while (0)
    int lpr = h->spmm_lpr;
    if (lpr != 8 && lpr != 16 && lpr != 32) {
        // measured on B200 (tools/profile_spmm2.py): best when every lane carries about two double2 row-fragments
        // per stored block, i.e. D * CPL2 ~ 2
        const int want = npairs * D / 2;
        lpr = (want <= 8) ? 8 : (want <= 16 ? 16 : 32);
    }
    while (lpr < 32 && lpr * 4 < npairs) lpr *= 2;
    if (D > 4) lpr = 32;
    const int cpl = (npairs + lpr - 1) / lpr;
    if (lpr == 8) { if (cpl <= 1) RVGP_V2(8, 1); else if (cpl <= 2) RVGP_V2(8, 2); else RVGP_V2(8, 4); }
    else if (lpr == 16) { if (cpl <= 1) RVGP_V2(16, 1); else RVGP_V2(16, 2); }
    else { if (cpl <= 1) RVGP_V2(32, 1); else RVGP_V2(32, 2); }
#undef RVGP_V2
    RVGP_LAUNCH_OK(h, "bsr_spmm_v2_kernel");
    return RVGP_OK;
}

template <int D, bool PATTERN>
static int launch_spmm_d(Handle* h, int nbrows, const int* indptr, const int* indices, const double* vals,
                         const double* X, int64_t ldx, const double* W, int64_t ldw, double* Y, int64_t ldy,
                         int ncols, double alpha, double beta, double gamma) {
    const int rpg = 8;  // rows per lane-group per CTA -> 64 * (32/LPR) consecutive block rows per CTA
#define RVGP_SPMM_LAUNCH(LPR, CPL)                                                                     \
    do {                                                                                               \
        constexpr int U = (D * CPL <= 2) ? 4 : ((D * CPL <= 8) ? 2 : 1);                               \
        const int rows_per_cta = 8 * (32 / LPR) * rpg;                                                 \
        const int grid = cdiv(nbrows, rows_per_cta);                                                   \
        bsr_spmm_kernel<D, LPR, CPL, U, PATTERN><<<grid, 256, 0, h->stream>>>(                         \
            nbrows, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma, rpg);    \
    } while (0)
    // ncu (profiles/r01): a warp-wide 64-bit load costs 2 L1 data-pipe wavefronts even when every lane reads the same
    // address, so the per-entry block values dominate when 32 lanes share one block row.  With LPR = 8 four block rows
    // share a warp: value / index loads are amortised over four entries and each lane carries ncols/8 columns.
    int lpr = h->spmm_lpr;
    if (lpr != 8 && lpr != 16 && lpr != 32) lpr = (ncols <= 8 * 8 && D <= 4) ? 8 : 32;
    while (lpr < 32 && lpr * 8 < ncols) lpr *= 2;
    if (D > 4 && lpr < 32 && ncols > lpr * 2) lpr = (ncols <= 32) ? 16 : 32;     // register budget
    const int cpl = (ncols + lpr - 1) / lpr;
    if (lpr == 8) {
        if (cpl <= 1) RVGP_SPMM_LAUNCH(8, 1); else if (cpl <= 2) RVGP_SPMM_LAUNCH(8, 2);
        else if (cpl <= 4) RVGP_SPMM_LAUNCH(8, 4); else RVGP_SPMM_LAUNCH(8, 8);
    } else if (lpr == 16) {
        if (cpl <= 1) RVGP_SPMM_LAUNCH(16, 1); else if (cpl <= 2) RVGP_SPMM_LAUNCH(16, 2);
        else RVGP_SPMM_LAUNCH(16, 4);
    } else {
        if (cpl <= 1) RVGP_SPMM_LAUNCH(32, 1); else RVGP_SPMM_LAUNCH(32, 2);
    }
#undef RVGP_SPMM_LAUNCH
    RVGP_LAUNCH_OK(h, "bsr_spmm_kernel");
    return RVGP_OK;
}

static int spmm_dispatch(Handle* h, int nbrows, int d, const int* indptr, const int* indices, const double* vals,
                         const double* X, int64_t ldx, const double* W, int64_t ldw, double* Y, int64_t ldy,
                         int ncols, double alpha, double beta, double gamma) {
    RVGP_REQUIRE(h, nbrows >= 0 && ncols >= 1 && ncols <= 128, "spmm: ncols must be in [1,128]");
    RVGP_REQUIRE(h, d == -2 || d >= 1, "spmm: bad block size");
    RVGP_REQUIRE(h, gamma == 0.0 || W != nullptr, "spmm: W required when gamma != 0");
    RVGP_REQUIRE(h, Y != X && Y != W, "spmm: Y must not alias X or W");
    if (nbrows == 0) return RVGP_OK;
    const bool aligned0 = (ncols % 2 == 0) && (ldx % 2 == 0) && (ldy % 2 == 0) && (W == nullptr || ldw % 2 == 0) &&
                         ((uintptr_t)X % 16 == 0) && ((uintptr_t)Y % 16 == 0) && ((uintptr_t)W % 16 == 0) &&
                         (vals == nullptr || (uintptr_t)vals % 16 == 0);
    RVGP_REQUIRE(h, ncols <= 64 || aligned0, "spmm: more than 64 columns needs the 128-bit path (even ncols / ld, 16-byte alignment)");
    const bool aligned = aligned0 && (!h->spmm_v1 || ncols > 64);
    RVGP_REQUIRE(h, h->rowlist == nullptr || aligned, "spmm: a row list needs the 128-bit path (even ncols / ld, 16-byte alignment)");
    if (d == -2) {   // ROT2-compressed 2x2 blocks (see rvgp_bsr_compress_rot2)
        RVGP_REQUIRE(h, aligned && vals != nullptr, "spmm: rot2 storage needs even ncols / leading dimensions and 16-byte aligned buffers");
        return launch_spmm_v2<2, false, true>(h, nbrows, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma);
    }
    if (vals == nullptr) {
        RVGP_REQUIRE(h, d == 1, "spmm: pattern mode (vals == NULL) needs d == 1");
        if (aligned) return launch_spmm_v2<1, true, false>(h, nbrows, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma);
        return launch_spmm_d<1, true>(h, nbrows, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma);
    }
    if (aligned) {
        switch (d) {
#define RVGP_CASE(DD) case DD: return launch_spmm_v2<DD, false, false>(h, nbrows, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma);
            RVGP_CASE(1) RVGP_CASE(2) RVGP_CASE(3) RVGP_CASE(4) RVGP_CASE(5) RVGP_CASE(6) RVGP_CASE(7) RVGP_CASE(8)
#undef RVGP_CASE
            default: break;
        }
    }
    switch (d) {
#define RVGP_CASE(DD) case DD: return launch_spmm_d<DD, false>(h, nbrows, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma);
        RVGP_CASE(1) RVGP_CASE(2) RVGP_CASE(3) RVGP_CASE(4) RVGP_CASE(5) RVGP_CASE(6) RVGP_CASE(7) RVGP_CASE(8)
#undef RVGP_CASE
        default: return set_error(h, RVGP_ERR_BAD_ARG, "spmm: block size d must be in [1,8]%s%s");
    }
}

int spmm_dispatch_public(Handle* h, int nbrows, int d, const int* indptr, const int* indices, const double* vals,
                         const double* X, int64_t ldx, const double* W, int64_t ldw, double* Y, int64_t ldy, int ncols,
                         double alpha, double beta, double gamma) {
    return spmm_dispatch(h, nbrows, d, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma);
}

}  // namespace rvgp

using namespace rvgp;

// Same as rvgp_bsr_spmm_f64 restricted to the block rows listed in `rowlist` (nlist entries); other rows of Y are left
// untouched.  Used to recompute the boundary rows of a row-sharded matrix after the halo exchange has landed, while the
// full launch overlapped the exchange (rvgp_b200/distributed.py).
extern "C" int rvgp_bsr_spmm_rows_f64(rvgp_handle_t hh, int nbrows, int d, const int32_t* indptr, const int32_t* indices,
                                      const double* vals, const double* X, int64_t ldx, const double* W, int64_t ldw,
                                      double* Y, int64_t ldy, int ncols, double alpha, double beta, double gamma,
                                      const int32_t* rowlist, int nlist) {
    Handle* h = H(hh);
    if (nlist <= 0) return RVGP_OK;
    h->rowlist = rowlist;
    h->nlist = nlist;
    const int rc = spmm_dispatch(h, nbrows, d, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma);
    h->rowlist = nullptr;
    h->nlist = 0;
    return rc;
}

extern "C" int rvgp_bsr_spmm_f64(rvgp_handle_t hh, int nbrows, int d, const int32_t* indptr, const int32_t* indices,
                                 const double* vals, const double* X, int64_t ldx, const double* W, int64_t ldw,
                                 double* Y, int64_t ldy, int ncols, double alpha, double beta, double gamma) {
    Handle* h = H(hh);
    return spmm_dispatch(h, nbrows, d, indptr, indices, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma);
}

namespace rvgp {
__global__ void compress_rot2_kernel(int64_t nnzb, const double* __restrict__ vals, const int* __restrict__ indices,
                                     double2* __restrict__ ab, int* __restrict__ idx_flag, int* __restrict__ bad, double rtol) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnzb) return;
    const double m00 = vals[4 * e], m01 = vals[4 * e + 1], m10 = vals[4 * e + 2], m11 = vals[4 * e + 3];
    const double scale = fabs(m00) + fabs(m01) + fabs(m10) + fabs(m11);
    // det sign decides rotation (s=+1: m11 = a, m01 = -b) or reflection (s=-1: m11 = -a, m01 = b)
    const bool flip = (m00 * m11 - m01 * m10) < 0.0;
    const double sg = flip ? -1.0 : 1.0;
    if (fabs(m11 - sg * m00) > rtol * scale || fabs(m01 + sg * m10) > rtol * scale) atomicOr(bad, 1);
    ab[e] = make_double2(m00, m10);
    idx_flag[e] = indices[e] | (flip ? (int)0x80000000 : 0);
}
}  // namespace rvgp

// Compress 2x2 blocks that are scaled orthogonal matrices to (a, b) + a flip bit in the column index (see ROT2 above).
// bad_flag (device int32, zeroed here) gets bit0 set when some block is NOT of that form within rtol (then keep using
// the plain storage).  Pass the outputs to rvgp_bsr_spmm_f64 / rvgp_cheb_filter_f64 with d = -2.
extern "C" int rvgp_bsr_compress_rot2(rvgp_handle_t hh, int64_t nnzb, const double* vals, const int32_t* indices, double* ab,
                                      int32_t* idx_flag, int32_t* bad_flag, double rtol) {
    Handle* h = H(hh);
    RVGP_CUDA_OK(h, cudaMemsetAsync(bad_flag, 0, sizeof(int), h->stream));
    if (nnzb == 0) return RVGP_OK;
    compress_rot2_kernel<<<cdiv(nnzb, 256), 256, 0, h->stream>>>(nnzb, vals, indices, reinterpret_cast<double2*>(ab), idx_flag,
                                                                  bad_flag, rtol);
    RVGP_LAUNCH_OK(h, "compress_rot2_kernel");
    return RVGP_OK;
}

// Scaled Chebyshev filter (Zhou & Saad, "A Chebyshev-Davidson algorithm", Alg. 3.1 form):
//   e = (hi - lo_cut)/2, c = (hi + lo_cut)/2, sigma_1 = e / (lo_spec - c), tau = 2 / sigma_1
//   Y_1     = (A V - c V) * sigma_1 / e
//   Y_{i+1} = 2 sigma_{i+1}/e (A Y_i - c Y_i) - sigma_i sigma_{i+1} Y_{i-1},   sigma_{i+1} = 1/(tau - sigma_i)
// Every step is ONE fused SpMM launch.  The three block vectors rotate; the result is copied back into V
// when it does not land there.
template <class Apply>
static int cheb_recurrence(Handle* h, Apply&& apply, int64_t nrows, double* buf[3], int64_t ld[3], int ncols, int degree,
                           double lo_spec, double lo_cut, double hi, int* result_slot) {
    const double e = 0.5 * (hi - lo_cut), c = 0.5 * (hi + lo_cut);
    const double sigma1 = e / (lo_spec - c), tau = 2.0 / sigma1;
    double sigma = sigma1;
    int rc = apply(buf[0], ld[0], (const double*)nullptr, (int64_t)0, buf[1], ld[1], sigma1 / e, -c * sigma1 / e, 0.0);
    if (rc) return rc;
    int prev = 0, cur = 1;
    for (int i = 2; i <= degree; ++i) {
        const double sn = 1.0 / (tau - sigma);
        const int nxt = 3 - prev - cur;
        rc = apply(buf[cur], ld[cur], (const double*)buf[prev], ld[prev], buf[nxt], ld[nxt], 2.0 * sn / e, -2.0 * sn * c / e,
                   -sigma * sn);
        if (rc) return rc;
        sigma = sn;
        prev = cur;
        cur = nxt;
    }
    *result_slot = cur;
    return RVGP_OK;
}

extern "C" int rvgp_cheb_filter_f64(rvgp_handle_t hh, int nbrows, int d, const int32_t* indptr, const int32_t* indices,
                                    const double* vals, double* V, int64_t ldv, double* work0, double* work1,
                                    int64_t ldw, int ncols, int degree, double lo_spec, double lo_cut, double hi) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, degree >= 0, "cheb_filter: degree >= 0");
    RVGP_REQUIRE(h, hi > lo_cut && lo_cut > lo_spec, "cheb_filter: need lo_spec < lo_cut < hi");
    if (degree == 0 || nbrows == 0) return RVGP_OK;
    double* buf[3] = {V, work0, work1};
    int64_t ld[3] = {ldv, ldw, ldw};
    int slot = 0;
    auto apply = [&](const double* X, int64_t ldx, const double* W, int64_t ldw_, double* Y, int64_t ldy, double a, double b,
                     double g) { return spmm_dispatch(h, nbrows, d, indptr, indices, vals, X, ldx, W, ldw_, Y, ldy, ncols, a, b, g); };
    int rc = cheb_recurrence(h, apply, (int64_t)nbrows * (d < 0 ? -d : d), buf, ld, ncols, degree, lo_spec, lo_cut, hi, &slot);
    if (rc) return rc;
    if (slot != 0) {
        RVGP_CUDA_OK(h, cudaMemcpy2DAsync(V, ldv * sizeof(double), buf[slot], ld[slot] * sizeof(double),
                                          (size_t)ncols * sizeof(double), (size_t)nbrows * (d < 0 ? -d : d),
                                          cudaMemcpyDeviceToDevice, h->stream));
    }
    return RVGP_OK;
}

namespace rvgp {
int spmm_mma_dispatch(Handle* h, int nbrows, int d, const int* kptr, const int* kcols, const double* afrag,
                      const double* X, int64_t ldx, const double* W, int64_t ldw, double* Y, int64_t ldy, int ncols,
                      double alpha, double beta, double gamma);
int spmm_mma_native_dispatch(Handle* h, int nbrows, const int* kptr, const int* kcols, const double* afrag, int rotc,
                             const double* X, int64_t nsx, const double* W, int64_t nsw, double* Y, int64_t nsy,
                             int ncols, double alpha, double beta, double gamma, int reverse);
int native_convert(Handle* h, bool to_native, int nbrows, int ncols, double* V, int64_t ldv, double* Xn, int64_t ns);
}

// Same filter through the FP64-MMA row-group SpMM (spmm_mma.cu).
//   d == 2 and work2 != NULL: the recurrence runs on NODE-CONTIGUOUS panels (rvgp_bsr_spmm_mma_native_f64): V is converted
//     into work0, the three (nrows x ncols, contiguous) work panels rotate, the result is converted back into V.
//     rotc != 0: kcols / afrag are the compact plan of rvgp_bsr_mma_rotc.
//   otherwise: row-major kernel (rvgp_bsr_spmm_mma_f64), work2 unused.
extern "C" int rvgp_cheb_filter_mma_f64(rvgp_handle_t hh, int nbrows, int d, const int32_t* kptr, const int32_t* kcols,
                                        const double* afrag, int rotc, double* V, int64_t ldv, double* work0, double* work1,
                                        double* work2, int64_t ldw, int ncols, int degree, double lo_spec, double lo_cut,
                                        double hi) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, degree >= 0, "cheb_filter: degree >= 0");
    RVGP_REQUIRE(h, hi > lo_cut && lo_cut > lo_spec, "cheb_filter: need lo_spec < lo_cut < hi");
    if (degree == 0 || nbrows == 0) return RVGP_OK;
    int slot = 0;
    if (d == 1 && rotc == 2) {
        // scalar unit-weight Laplacian as L (x) I_2 (spmm_mma.cu, AMODE 2): the ROW-MAJOR panels are already in the native
        // layout with ncols / 2 columns, so V itself is the first buffer and nothing is converted
        RVGP_REQUIRE(h, ncols % 32 == 0 && ldv % 2 == 0 && ldw % 2 == 0, "cheb_filter_mma (pattern): ncols % 32 == 0 and even leading dimensions");
        double* buf[3] = {V, work0, work1};
        int64_t ld[3] = {ldv, ldw, ldw};
        auto apply = [&](const double* X, int64_t nsx, const double* W, int64_t nsw, double* Y, int64_t nsy, double a, double b,
                         double g) {
            return spmm_mma_native_dispatch(h, nbrows, kptr, kcols, afrag, 2, X, nsx, W, nsw, Y, nsy, ncols / 2, a, b, g, 0);
        };
        int rc = cheb_recurrence(h, apply, (int64_t)nbrows, buf, ld, ncols, degree, lo_spec, lo_cut, hi, &slot);
        if (rc) return rc;
        if (slot != 0) {
            RVGP_CUDA_OK(h, cudaMemcpy2DAsync(V, ldv * sizeof(double), buf[slot], ld[slot] * sizeof(double),
                                              (size_t)ncols * sizeof(double), (size_t)nbrows, cudaMemcpyDeviceToDevice, h->stream));
        }
        return RVGP_OK;
    }
    if (d == 2 && work2 != nullptr) {
        RVGP_REQUIRE(h, ldw == ncols, "cheb_filter_mma: the native path needs contiguous work panels (ldw == ncols)");
        const int64_t ns = 2 * (int64_t)ncols;
        int rc = native_convert(h, true, nbrows, ncols, V, ldv, work0, ns);
        if (rc) return rc;
        double* buf[3] = {work0, work1, work2};
        int64_t ld[3] = {ns, ns, ns};
        auto apply = [&](const double* X, int64_t nsx, const double* W, int64_t nsw, double* Y, int64_t nsy, double a, double b,
                         double g) {
            return spmm_mma_native_dispatch(h, nbrows, kptr, kcols, afrag, rotc, X, nsx, W, nsw, Y, nsy, ncols, a, b, g, 0);
        };
        rc = cheb_recurrence(h, apply, (int64_t)nbrows * d, buf, ld, ncols, degree, lo_spec, lo_cut, hi, &slot);
        if (rc) return rc;
        return native_convert(h, false, nbrows, ncols, V, ldv, buf[slot], ns);
    }
    RVGP_REQUIRE(h, rotc == 0, "cheb_filter_mma: the compact plan needs the native path (d == 2, work2 != NULL)");
    double* buf[3] = {V, work0, work1};
    int64_t ld[3] = {ldv, ldw, ldw};
    auto apply = [&](const double* X, int64_t ldx, const double* W, int64_t ldw_, double* Y, int64_t ldy, double a, double b,
                     double g) { return spmm_mma_dispatch(h, nbrows, d, kptr, kcols, afrag, X, ldx, W, ldw_, Y, ldy, ncols, a, b, g); };
    int rc = cheb_recurrence(h, apply, (int64_t)nbrows * d, buf, ld, ncols, degree, lo_spec, lo_cut, hi, &slot);
    if (rc) return rc;
    if (slot != 0) {
        RVGP_CUDA_OK(h, cudaMemcpy2DAsync(V, ldv * sizeof(double), buf[slot], ld[slot] * sizeof(double),
                                          (size_t)ncols * sizeof(double), (size_t)nbrows * d, cudaMemcpyDeviceToDevice,
                                          h->stream));
    }
    return RVGP_OK;
}

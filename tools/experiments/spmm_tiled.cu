// K9 (v2): block-CSR SpMM with tile-local gather staged through shared memory (cp.async).
//
// ncu on the gather kernel (spmm.cu; profiles/r01_*) shows DRAM traffic == algorithmic bytes but the L1
// wavefront pipe at 82 %: every neighbour row is fetched through L1 ~12 times and every block value / index is a
// separate warp-uniform load.  This version removes both from the L1 path:
//   * rows are processed in tiles of TR consecutive (Morton-ordered) block rows; a one-off PLAN lists each tile's
//     unique neighbour nodes and rewrites every entry's column as a 16-bit index into that list;
//   * one CTA per tile: the unique neighbours' rows of X (and the tile's matrix values) are copied global -> shared
//     by 16-byte cp.async (LDGSTS), so they cross L2 -> SM once per tile instead of once per use; tiles with more
//     than `umax` unique neighbours (Morton-curve jumps; a few %) gather from global memory instead;
//   * the compute loop then reads X rows with conflict-free LDS.64, block values with LDS.128 and the local
//     indices with one LDS per 32 entries + warp shuffles.
// Several CTAs are resident per SM, so one tile's staging overlaps another's compute.
// Y = alpha * (A @ X) + beta * X + gamma * W, same contract as rvgp_bsr_spmm_f64.
#include "common.cuh"

namespace rvgp {

constexpr int PLAN_ECAP = 2048;   // max stored entries per tile

// ---- plan ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_plan_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, int n, int TR, int ucap,
                 int* __restrict__ tile_u, int* __restrict__ ucols, unsigned short* __restrict__ lidx,
                 int* __restrict__ flags /* [0] max U, [1] invalid, [2] max entries per tile */) {
    __shared__ int cols[PLAN_ECAP];
    __shared__ int pos[PLAN_ECAP];
    __shared__ int uniq[PLAN_ECAP];
    const int t = blockIdx.x, tid = threadIdx.x;
    const int r0 = t * TR, r1 = min(n, r0 + TR);
    const int e0 = indptr[r0], e1 = indptr[r1], ne = e1 - e0;
    if (ne > PLAN_ECAP) { if (tid == 0) atomicOr(flags + 1, 1); return; }
    int P = 1;
    while (P < ne) P <<= 1;
    for (int i = tid; i < P; i += 256) cols[i] = (i < ne) ? indices[e0 + i] : 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += 256) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const int a = cols[i], b = cols[ixj];
                    const bool up = ((i & k) == 0);
                    if ((a > b) == up) { cols[i] = b; cols[ixj] = a; }
                }
            }
            __syncthreads();
        }
    // inclusive scan of "head" flags (Hillis-Steele, double-buffered through pos/uniq)
    for (int i = tid; i < P; i += 256) pos[i] = (i < ne && (i == 0 || cols[i] != cols[i - 1])) ? 1 : 0;
    __syncthreads();
    int* src = pos; int* dst = uniq;
    for (int off = 1; off < P; off <<= 1) {
        for (int i = tid; i < P; i += 256) dst[i] = src[i] + ((i >= off) ? src[i - off] : 0);
        __syncthreads();
        int* tmp = src; src = dst; dst = tmp;
    }
    const int U = (ne > 0) ? src[ne - 1] : 0;
    if (U > ucap) { if (tid == 0) atomicOr(flags + 1, 1); return; }
    // scatter the unique values (dst is free now)
    for (int i = tid; i < ne; i += 256)
        if (i == 0 || cols[i] != cols[i - 1]) dst[src[i] - 1] = cols[i];
    __syncthreads();
    for (int i = tid; i < U; i += 256) ucols[(int64_t)t * ucap + i] = dst[i];
    if (tid == 0) { tile_u[t] = U; atomicMax(flags, U); atomicMax(flags + 2, ne); }
    for (int i = tid; i < ne; i += 256) {
        const int c = indices[e0 + i];
        int lo = 0, hi = U - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (dst[mid] < c) lo = mid + 1; else hi = mid; }
        lidx[e0 + i] = (unsigned short)lo;
    }
}

// ---- PTX helpers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- SpMM -------------------------------------------------------------------------------------------------------
// dynamic smem: [bar 16 B][Xs: U * D * ncols doubles][Vs: ne * D*D doubles][Ls: ne u16][Us: U int (pattern mode)]
template <int D, int CPL, bool PATTERN>
__global__ void __launch_bounds__(256)
bsr_spmm_tiled_kernel(int nbrows, int TR, int ucap, const int* __restrict__ indptr, const int* __restrict__ tile_u,
                      const int* __restrict__ ucols, const unsigned short* __restrict__ lidx,
                      const double* __restrict__ vals, const double* __restrict__ X, int64_t ldx,
                      const double* __restrict__ W, int64_t ldw, double* __restrict__ Y, int64_t ldy, int ncols,
                      double alpha, double beta, double gamma, int umax, int nemax) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Xs = reinterpret_cast<double*>(smem_raw + 16);
    double* Vs = Xs + (size_t)umax * D * ncols;
    unsigned short* Ls = reinterpret_cast<unsigned short*>(Vs + (PATTERN ? 0 : (size_t)nemax * D * D));
    int* Us = reinterpret_cast<int*>(Ls + ((nemax + 7) & ~7));     // ucap ints (pattern mode / heavy tiles)

    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = t * TR, r1 = min(nbrows, r0 + TR);
    const int te0 = __ldg(indptr + r0), te1 = __ldg(indptr + r1), ne = te1 - te0;
    const int U = __ldg(tile_u + t);
    const int* uc = ucols + (int64_t)t * ucap;
    const bool heavy = U > umax;                   // rare tile with too many unique neighbours: gather from global

    // stage the unique neighbours' rows of X (and the tile's block values) with 16-byte cp.async (LDGSTS, L2 -> smem,
    // no register staging); ~U*D*ncols/2 chunks spread over the 256 threads.  (A first version used one
    // cp.async.bulk per neighbour: correct, but 512-byte bulk copies are issue-bound in the TMA unit -- 2-3x slower.)
    {
        const int cpr = ncols >> 1;                       // 16-byte chunks per staged row
        const int nrow = heavy ? 0 : U * D;
        for (int c = tid; c < nrow * cpr; c += 256) {
            const int row = c / cpr, off = c - row * cpr;
            const int u = row / D, q = row - u * D;
            const double* src = X + ((int64_t)__ldg(uc + u) * D + q) * ldx + 2 * off;
            cp_async16(Xs + (size_t)row * ncols + 2 * off, src);
        }
        if (!PATTERN && ((D * D) % 2 == 0)) {
            const int nch = ne * ((D * D) / 2);
            const double* vsrc = vals + (int64_t)te0 * D * D;
            for (int c = tid; c < nch; c += 256) cp_async16(Vs + 2 * c, vsrc + 2 * c);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // meanwhile: local indices (and odd-sized value blocks) through ordinary loads
    for (int i = tid; i < ne; i += 256) Ls[i] = lidx[te0 + i];
    if (PATTERN || heavy)
        for (int i = tid; i < U; i += 256) Us[i] = __ldg(uc + i);
    if (!PATTERN && ((D * D) % 2 != 0))
        for (int i = tid; i < ne * D * D; i += 256) Vs[i] = __ldg(vals + (int64_t)te0 * D * D + i);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    bool colok[CPL];
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) colok[cc] = (lane + 32 * cc) < ncols;

    for (int i = r0 + warp; i < r1; i += 8) {
        const int e0 = __ldg(indptr + i) - te0, e1 = __ldg(indptr + i + 1) - te0;
        double acc[D][CPL];
#pragma unroll
        for (int p = 0; p < D; ++p)
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) acc[p][cc] = 0.0;
        for (int eb = e0; eb < e1; eb += 32) {
            const int cnt = min(32, e1 - eb);
            const int myl = (lane < cnt) ? (int)Ls[eb + lane] : 0;
#pragma unroll 2
            for (int u = 0; u < cnt; ++u) {
                const int li = __shfl_sync(0xffffffffu, myl, u);
                double x[D][CPL];
                if (!heavy) {
                    const double* xr = Xs + (size_t)li * D * ncols;
#pragma unroll
                    for (int q = 0; q < D; ++q)
#pragma unroll
                        for (int cc = 0; cc < CPL; ++cc) x[q][cc] = colok[cc] ? xr[q * ncols + lane + 32 * cc] : 0.0;
                } else {
                    const double* xr = X + (int64_t)Us[li] * D * ldx;
#pragma unroll
                    for (int q = 0; q < D; ++q)
#pragma unroll
                        for (int cc = 0; cc < CPL; ++cc) x[q][cc] = colok[cc] ? __ldg(xr + q * ldx + lane + 32 * cc) : 0.0;
                }
                if (PATTERN) {
                    const double r = (Us[li] == i) ? (double)(e1 - e0 - 1) : -1.0;
#pragma unroll
                    for (int cc = 0; cc < CPL; ++cc) acc[0][cc] = fma(r, x[0][cc], acc[0][cc]);
                } else {
                    const double* rp = Vs + (size_t)(eb + u) * (D * D);
                    double r[D * D];
                    if ((D * D) % 2 == 0) {
#pragma unroll
                        for (int v = 0; v < D * D; v += 2) {
                            const double2 rv = *reinterpret_cast<const double2*>(rp + v);
                            r[v] = rv.x; r[v + 1] = rv.y;
                        }
                    } else {
#pragma unroll
                        for (int v = 0; v < D * D; ++v) r[v] = rp[v];
                    }
#pragma unroll
                    for (int p = 0; p < D; ++p)
#pragma unroll
                        for (int q = 0; q < D; ++q)
#pragma unroll
                            for (int cc = 0; cc < CPL; ++cc) acc[p][cc] = fma(r[p * D + q], x[q][cc], acc[p][cc]);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < D; ++p)
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                if (!colok[cc]) continue;
                const int c = lane + 32 * cc;
                const int64_t r = (int64_t)i * D + p;
                double y = alpha * acc[p][cc];
                if (beta != 0.0) y = fma(beta, __ldg(X + r * ldx + c), y);
                if (gamma != 0.0) y = fma(gamma, __ldg(W + r * ldw + c), y);
                Y[r * ldy + c] = y;
            }
    }
}

template <int D, bool PATTERN>
static int launch_tiled(Handle* h, int nbrows, int TR, int ucap, int umax, int nemax, const int* indptr, const int* tile_u,
                        const int* ucols, const unsigned short* lidx, const double* vals, const double* X, int64_t ldx,
                        const double* W, int64_t ldw, double* Y, int64_t ldy, int ncols, double alpha, double beta,
                        double gamma) {
    const size_t smem = 16 + (size_t)umax * D * ncols * 8 + (PATTERN ? 0 : (size_t)nemax * D * D * 8) +
                        (size_t)((nemax + 7) & ~7) * 2 + (size_t)ucap * 4 + 16;
    if (smem > 200 * 1024) return set_error(h, RVGP_ERR_CAPACITY, "spmm_tiled: tile does not fit in shared memory%s%s");
    const int ntiles = cdiv(nbrows, TR);
#define RVGP_TILED(CPL)                                                                                              \
    do {                                                                                                             \
        auto kern = bsr_spmm_tiled_kernel<D, CPL, PATTERN>;                                                          \
        RVGP_CUDA_OK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
        kern<<<ntiles, 256, smem, h->stream>>>(nbrows, TR, ucap, indptr, tile_u, ucols, lidx, vals, X, ldx, W, ldw, \
                                               Y, ldy, ncols, alpha, beta, gamma, umax, nemax);                      \
    } while (0)
    if (ncols <= 32) RVGP_TILED(1);
    else RVGP_TILED(2);
#undef RVGP_TILED
    RVGP_LAUNCH_OK(h, "bsr_spmm_tiled_kernel");
    return RVGP_OK;
}

int spmm_tiled_dispatch(Handle* h, int nbrows, int d, int TR, int ucap, int umax, int nemax, const int* indptr,
                        const int* tile_u, const int* ucols, const unsigned short* lidx, const double* vals,
                        const double* X, int64_t ldx, const double* W, int64_t ldw, double* Y, int64_t ldy, int ncols,
                        double alpha, double beta, double gamma) {
    RVGP_REQUIRE(h, ncols >= 2 && ncols <= 64 && (ncols % 2) == 0, "spmm_tiled: ncols must be even and in [2,64]");
    RVGP_REQUIRE(h, (ldx % 2) == 0 && ((uintptr_t)X % 16) == 0, "spmm_tiled: X must be 16-byte aligned with even ldx");
    RVGP_REQUIRE(h, gamma == 0.0 || W != nullptr, "spmm_tiled: W required when gamma != 0");
    RVGP_REQUIRE(h, Y != X && Y != W, "spmm_tiled: Y must not alias X or W");
    if (nbrows == 0) return RVGP_OK;
    if (vals == nullptr) {
        RVGP_REQUIRE(h, d == 1, "spmm_tiled: pattern mode needs d == 1");
        return launch_tiled<1, true>(h, nbrows, TR, ucap, umax, nemax, indptr, tile_u, ucols, lidx, vals, X, ldx, W, ldw, Y,
                                     ldy, ncols, alpha, beta, gamma);
    }
    RVGP_REQUIRE(h, ((uintptr_t)vals % 16) == 0, "spmm_tiled: vals must be 16-byte aligned");
    switch (d) {
#define RVGP_CASE(DD) case DD: return launch_tiled<DD, false>(h, nbrows, TR, ucap, umax, nemax, indptr, tile_u, ucols, lidx, vals, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma);
        RVGP_CASE(1) RVGP_CASE(2) RVGP_CASE(3) RVGP_CASE(4) RVGP_CASE(5) RVGP_CASE(6) RVGP_CASE(7) RVGP_CASE(8)
#undef RVGP_CASE
        default: return set_error(h, RVGP_ERR_BAD_ARG, "spmm_tiled: block size d must be in [1,8]%s%s");
    }
}

}  // namespace rvgp

using namespace rvgp;

// Build the tile plan of a CSR pattern.  tile_u (ntiles), ucols (ntiles*ucap), lidx (nnzb, uint16);
// info (device int32[3], zeroed here): [0] max unique columns per tile, [1] != 0 -> plan invalid (a tile has more than
// ucap unique columns or more than 2048 entries; callers then keep using rvgp_bsr_spmm_f64), [2] max entries per tile.
extern "C" int rvgp_bsr_tile_plan(rvgp_handle_t hh, int nbrows, const int32_t* indptr, const int32_t* indices, int TR,
                                  int ucap, int32_t* tile_u, int32_t* ucols, uint16_t* lidx, int32_t* info) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, TR >= 8 && TR <= 256 && (TR % 8) == 0 && ucap >= 1 && ucap <= 65535, "tile_plan: bad TR / ucap");
    RVGP_CUDA_OK(h, cudaMemsetAsync(info, 0, 3 * sizeof(int), h->stream));
    if (nbrows == 0) return RVGP_OK;
    tile_plan_kernel<<<cdiv(nbrows, TR), 256, 0, h->stream>>>(indptr, indices, nbrows, TR, ucap, tile_u, ucols, lidx, info);
    RVGP_LAUNCH_OK(h, "tile_plan_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_bsr_spmm_tiled_f64(rvgp_handle_t hh, int nbrows, int d, int TR, int ucap, int umax, int nemax,
                                       const int32_t* indptr, const int32_t* tile_u, const int32_t* ucols,
                                       const uint16_t* lidx, const double* vals, const double* X, int64_t ldx,
                                       const double* W, int64_t ldw, double* Y, int64_t ldy, int ncols, double alpha,
                                       double beta, double gamma) {
    return spmm_tiled_dispatch(H(hh), nbrows, d, TR, ucap, umax, nemax, indptr, tile_u, ucols, lidx, vals, X, ldx, W, ldw, Y,
                               ldy, ncols, alpha, beta, gamma);
}

// K9 (v4): row-group block-SpMM on the FP64 tensor path (mma.sync.m8n8k4.f64).
//
//   Y = alpha * (A @ X) + beta * X + gamma * W        (same contract as rvgp_bsr_spmm_f64)
//
// Why: ncu on the gather kernel (profiles/r01_spmm_v2_Lc_b64_summary.txt) shows the L1 data pipe at 76 % with DRAM at
// 42 %.  The pipe returns 4 bytes per lane per wavefront, whatever the address pattern, so per stored 2x2 block the
// kernel pays 8 wavefronts for the two gathered X rows AND 9 more for the warp-uniform value / index loads (the same
// 36 bytes replicated into all 32 lanes): 235 wavefronts per node against a budget of ~250 cycles per node at 0.6 of
// the HBM roofline.  A register-tile FMA kernel cannot avoid the replication; an MMA can, because its A operand is
// DISTRIBUTED over the warp (one double per lane):
//   * G = 8/d consecutive (Morton-ordered) nodes form a row group = the 8 rows (M) of the MMA;
//   * the union of their neighbour lists is walked in "k-steps" of 4/d neighbour nodes = the 4 k-indices (K);
//   * per k-step the plan stores the dense 8x4 slice of A in fragment order (32 doubles, zeros where a node does not
//     store that neighbour): ONE coalesced 256-byte load per k-step instead of 8 uniform loads per stored block;
//   * the N dimension runs over the columns of the block vector: one LDG.128 per lane gathers 16 columns x 4 X-rows
//     (four fully used 128-byte lines) and feeds two MMAs (even / odd columns), so every gathered X row is reused by
//     all nodes of the group that store it (1.9x fewer gathered bytes at G = 4) and the accumulator fragment of a lane
//     is 4 CONSECUTIVE columns of one row, i.e. the epilogue is one 256-bit load / store per stream and chunk.
// Wavefronts per node drop from ~235 to ~85; the price is ~2x the FP64 work (zero fill), on a pipe that was 19 % busy.
// Needs d in {1, 2}, ncols % 16 == 0, 32-byte aligned X / W / Y with leading dimensions % 4 == 0 (the dispatcher in
// eigensolver.py falls back to the gather kernel otherwise).
#include <algorithm>

#include "common.cuh"

namespace rvgp {

__device__ __forceinline__ void mma_f64(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct double4v { double a, b, c, d; };

__device__ __forceinline__ double4v ld256_nc(const double* p) {
    double4v v;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p));
    return v;
}
__device__ __forceinline__ double4v ld256_stream(const double* p) {
    double4v v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p));
    return v;
}
__device__ __forceinline__ void st256(double* p, const double4v& v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "d"(v.d) : "memory");
}
__device__ __forceinline__ void st256_stream(double* p, const double4v& v) {
    asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "d"(v.d)
                 : "memory");
}

// fire-and-forget L2 prefetch of a contiguous byte range (16-byte aligned, size % 16 == 0): no registers held in flight
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l2_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// ---- plan: pack the k-step fragments ------------------------------------------------------------------------------
// One warp per row group.  gptr / uent come from rvgp_bsr_merge_plan (R = 8/d); kptr[g] = first k-step of group g
// (exclusive prefix of ceil(ulen_g / T), T = 4/d).  Outputs: kcols (T column indices per k-step; padding repeats the
// last real column with bit 31 set, its A entries are zero) and afrag (32 doubles per k-step, lane order: A[m = lane>>2][k = lane&3],
// m = local_node * d + p, k = local_neighbour * d + q).
template <int D>
__global__ void mma_pack_kernel(int nbrows, int ngroups, const int* __restrict__ indptr, const int* __restrict__ indices,
                                const double* __restrict__ vals, const int* __restrict__ gptr,
                                const int2* __restrict__ uent, const int* __restrict__ kptr, int* __restrict__ kcols,
                                double* __restrict__ afrag) {
    constexpr int R = 8 / D, T = 4 / D;
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= ngroups) return;
    const int m = lane >> 2, k = lane & 3;
    const int r = m / D, p = m % D, t = k / D, q = k % D;
    const int row = g * R + r;
    const int u0 = gptr[g], ulen = gptr[g + 1] - u0;
    const int s0 = kptr[g], ns = kptr[g + 1] - s0;
    int e0 = 0, e1 = 0;
    if (row < nbrows) { e0 = indptr[row]; e1 = indptr[row + 1]; }
    for (int s = 0; s < ns; ++s) {
        const int u = s * T + t;
        const int2 ent = uent[u0 + (u < ulen ? u : ulen - 1)];
        double v = 0.0;
        if (u < ulen && ((ent.y >> r) & 1)) {
            int lo = e0, hi = e1 - 1, e = -1;           // columns of a CSR row are sorted: binary search
            while (lo <= hi) {
                const int mid = (lo + hi) >> 1;
                const int c = indices[mid] & 0x7fffffff;
                if (c == ent.x) { e = mid; break; }
                if (c < ent.x) lo = mid + 1; else hi = mid - 1;
            }
            if (e >= 0) {
                if (vals) v = vals[(int64_t)e * (D * D) + p * D + q];
                else v = (ent.x == row) ? (double)(e1 - e0 - 1) : -1.0;     // unit-weight graph Laplacian (geometry.py:61)
            }
        }
        afrag[(int64_t)(s0 + s) * 32 + lane] = v;
        if (m == 0 && q == 0) kcols[(int64_t)(s0 + s) * T + t] = (u < ulen) ? ent.x : (ent.x | (int)0x80000000);   // bit 31 = padding
    }
}

// ---- SpMM ---------------------------------------------------------------------------------------------------------
// One warp per row group, NCH chunks of 16 columns per warp (blockIdx.y selects further column slabs of NCH*16).
template <int D, int NCH>
__global__ void __launch_bounds__(128, NCH == 4 ? 4 : 6)
bsr_spmm_mma_kernel(int ngroups, int64_t nrows, const int* __restrict__ kptr, const int* __restrict__ kcols,
                    const double* __restrict__ afrag, const double* __restrict__ X, int64_t ldx,
                    const double* __restrict__ W, int64_t ldw, double* __restrict__ Y, int64_t ldy, double alpha,
                    double beta, double gamma, int groups_per_warp, int stream_policy) {
    constexpr int T = 4 / D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kq = lane & 3, nn = lane >> 2;
    const int t = kq / D, q = kq % D;
    const int coloff = blockIdx.y * (NCH * 16);
    const double* Xg = X + coloff + 2 * nn;          // gather base: lane reads columns 16c + 2nn, 2nn+1 of X-row k
    const int g0 = blockIdx.x * (4 * groups_per_warp);

    for (int it = 0; it < groups_per_warp; ++it) {
        const int g = g0 + it * 4 + warp;
        if (g >= ngroups) break;
        const int s0 = __ldg(kptr + g), s1 = __ldg(kptr + g + 1);
        double acc[NCH][2][2];
#pragma unroll
        for (int c = 0; c < NCH; ++c) { acc[c][0][0] = acc[c][0][1] = acc[c][1][0] = acc[c][1][1] = 0.0; }

        if (s0 < s1) {
            int col = __ldg(kcols + (int64_t)s0 * T + t) & 0x7fffffff;
            int col_nxt = (s0 + 1 < s1) ? (__ldg(kcols + (int64_t)(s0 + 1) * T + t) & 0x7fffffff) : 0;
            double a = __ldg(afrag + (int64_t)s0 * 32 + lane);
            double2 xb[NCH];
            {
                const double* xr = Xg + ((int64_t)col * D + q) * ldx;
#pragma unroll
                for (int c = 0; c < NCH; ++c) xb[c] = __ldg(reinterpret_cast<const double2*>(xr + 16 * c));
            }
            for (int s = s0; s < s1; ++s) {
                double a_n = 0.0;
                double2 xn[NCH];
                int col_n2 = 0;
                if (s + 1 < s1) {                                  // warp-uniform: software pipeline, one k-step ahead
                    const double* xr = Xg + ((int64_t)col_nxt * D + q) * ldx;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) xn[c] = __ldg(reinterpret_cast<const double2*>(xr + 16 * c));
                    a_n = __ldg(afrag + (int64_t)(s + 1) * 32 + lane);
                    if (s + 2 < s1) col_n2 = __ldg(kcols + (int64_t)(s + 2) * T + t) & 0x7fffffff;
                }
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    mma_f64(acc[c][0][0], acc[c][0][1], a, xb[c].x);     // even columns of the chunk
                    mma_f64(acc[c][1][0], acc[c][1][1], a, xb[c].y);     // odd columns
                }
                if (s + 1 < s1) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) xb[c] = xn[c];
                    a = a_n;
                    col_nxt = col_n2;
                }
            }
        }
        // epilogue: lane holds row m = nn, columns 16c + 4kq .. +3 = {even0, odd0, even1, odd1}
        const int64_t row = (int64_t)g * 8 + nn;
        if (row < nrows) {
            const int64_t co = coloff + 4 * kq;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                double4v y;
                y.a = alpha * acc[c][0][0]; y.b = alpha * acc[c][1][0]; y.c = alpha * acc[c][0][1]; y.d = alpha * acc[c][1][1];
                if (beta != 0.0) {
                    const double4v xv = ld256_nc(X + row * ldx + co + 16 * c);
                    y.a = fma(beta, xv.a, y.a); y.b = fma(beta, xv.b, y.b); y.c = fma(beta, xv.c, y.c); y.d = fma(beta, xv.d, y.d);
                }
                if (gamma != 0.0) {
                    const double* wp = W + row * ldw + co + 16 * c;
                    const double4v wv = (stream_policy & 1) ? ld256_stream(wp) : ld256_nc(wp);
                    y.a = fma(gamma, wv.a, y.a); y.b = fma(gamma, wv.b, y.b); y.c = fma(gamma, wv.c, y.c); y.d = fma(gamma, wv.d, y.d);
                }
                double* yp = Y + row * ldy + co + 16 * c;
                if (stream_policy & 2) st256_stream(yp, y); else st256(yp, y);
            }
        }
    }
}

int spmm_mma_dispatch(Handle* h, int nbrows, int d, const int* kptr, const int* kcols, const double* afrag,
                      const double* X, int64_t ldx, const double* W, int64_t ldw, double* Y, int64_t ldy, int ncols,
                      double alpha, double beta, double gamma) {
    RVGP_REQUIRE(h, d == 1 || d == 2, "spmm_mma: block size d must be 1 or 2");
    RVGP_REQUIRE(h, ncols >= 16 && ncols % 16 == 0, "spmm_mma: ncols must be a multiple of 16");
    RVGP_REQUIRE(h, ldx % 4 == 0 && ldy % 4 == 0 && (W == nullptr || ldw % 4 == 0) && (uintptr_t)X % 32 == 0 &&
                        (uintptr_t)Y % 32 == 0 && (uintptr_t)W % 32 == 0,
                 "spmm_mma: buffers must be 32-byte aligned with leading dimensions divisible by 4");
    RVGP_REQUIRE(h, gamma == 0.0 || W != nullptr, "spmm_mma: W required when gamma != 0");
    RVGP_REQUIRE(h, Y != X && Y != W, "spmm_mma: Y must not alias X or W");
    if (nbrows == 0) return RVGP_OK;
    const int R = 8 / d;
    const int ngroups = cdiv(nbrows, R);
    const int64_t nrows = (int64_t)nbrows * d;
    const int gpw = h->mma_gpw > 0 ? h->mma_gpw : 8;
    const int nch = (ncols % 64 == 0) ? 4 : ((ncols % 32 == 0) ? 2 : 1);
    dim3 grid(cdiv(ngroups, 4 * gpw), ncols / (16 * nch));
#define RVGP_MMA(DD, NCH)                                                                                        \
    bsr_spmm_mma_kernel<DD, NCH><<<grid, 128, 0, h->stream>>>(ngroups, nrows, kptr, kcols, afrag, X, ldx, W, ldw, Y, ldy, \
                                                              alpha, beta, gamma, gpw, h->mma_stream_policy)
    if (d == 2) { if (nch == 4) RVGP_MMA(2, 4); else if (nch == 2) RVGP_MMA(2, 2); else RVGP_MMA(2, 1); }
    else        { if (nch == 4) RVGP_MMA(1, 4); else if (nch == 2) RVGP_MMA(1, 2); else RVGP_MMA(1, 1); }
#undef RVGP_MMA
    RVGP_LAUNCH_OK(h, "bsr_spmm_mma_kernel");
    return RVGP_OK;
}

// ---- native-layout variant (d == 2) ---------------------------------------------------------------------------------
// ncu on the row-major kernel above (profiles/r01_spmm_mma_rowmajor_summary.txt): a quarter-warp of the B-fragment gather
// (lanes n0, n0+1 x k = 0..3) touches FOUR X rows = four 128-byte lines, so one LDG.128 costs 16 L1 wavefronts instead
// of 4 and the kernel is slower than the gather kernel.  The fragment order cannot change, the memory layout can: the
// Chebyshev recurrence only ever feeds its own outputs back in, so inside rvgp_cheb_filter_mma_f64 the three rotating
// panels live in a NODE-CONTIGUOUS layout
//      Xn[node][cp][q][e]      cp = column pair, q = component (0..1), e = column parity;  element (2*node+q, 2*cp+e)
// Now the two components of a node sit in the same line: 2 lines per quarter-warp, 8 wavefronts per gather.  The beta * X
// term is folded into the A fragment of the diagonal entries (a += beta / alpha), which removes the X-own stream.
//
// ROTC: every 2x2 block of the connection Laplacian is a scaled rotation / reflection [[a, -s b], [b, s a]] (Procrustes
// factor U V^T, diagonal deg * I), so a k-step needs 16 doubles (a, b per (node, neighbour)) + 8 sign bits instead of 32
// doubles; the signs ride in bits 27..30 of the two column words.  Halves the matrix stream (0.86 -> 0.43 GB at C4).
//
// Cache policy (policy bits): X gathers evict_last in L2 (every X row is gathered ~7x by neighbouring groups over a
// short time window), A fragments / W evict_first and not allocated in L1 (read exactly once).
__device__ __forceinline__ uint64_t make_policy(int kind) {        // 0 normal, 1 evict_last, 2 evict_first
    uint64_t pol;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double2 ldg128_hint(const double* p, uint64_t pol) {
    double2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double2 ldg128_stream_hint(const double* p, uint64_t pol) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ldg64_stream_hint(const double* p, uint64_t pol) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void stg128_hint(double* p, double2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" :: "l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void prefetch_l2_bulk_hint(const void* p, unsigned bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" :: "l"(p), "r"(bytes), "l"(pol) : "memory");
}

constexpr int ROTC_COLMASK = 0x07ffffff;     // column index bits of a ROTC column word (bits 27..30 = flips, bit 31 = padding)

// Schedule: `persistent` != 0 launches sm_count * MINB CTAs and strides the warps over the row groups, so that at any time
// the ~3000 warps of the chip work on ~3000 CONSECUTIVE groups (12 k nodes, ~40 MB of X / W / Y / A): every re-read of an
// X row by a neighbouring group is an L2 hit, and the NT/32 adjacent groups of a CTA share their gathers in L1.
// Otherwise CTA b owns the contiguous chunk of nw * groups_per_warp groups (the v2 schedule).
// gamma * W is loaded straight into the accumulators at group start (acc = (gamma/alpha) W), so there is no load in the epilogue.
// AMODE selects where the A fragment comes from: 0 = full 32-double fragments, 1 = ROTC (compact rotations, above),
// 2 = PATTERN: the unit-weight scalar graph Laplacian L (geometry.py:55-63) run as L (x) I_2 -- a ROW-MAJOR (n x B) panel of a
// scalar block vector IS a node-contiguous panel of the d = 2 layout with B/2 columns (element (node, 4cp + 2q + e) <-> pseudo
// component q of column 2cp + e), and every 2x2 block is a * I_2 with a in {deg_i, -1, 0}.  Nothing is streamed for the
// matrix except the column words: bits 27..30 of word t say which of the group's 4 nodes store that neighbour (the R-bit row
// mask of the merge plan), and the diagonal value deg_i comes from a 4-byte-per-node array (passed through `afrag`).
template <int NCH, int KS, int NT, int MINB, int AMODE>
__global__ void __launch_bounds__(NT, MINB)
bsr_spmm_mma_native_kernel(int ngroups, int nbrows, const int* __restrict__ kptr, const int* __restrict__ kcols,
                           const double* __restrict__ afrag, const double* __restrict__ X, int64_t nsx,
                           const double* __restrict__ W, int64_t nsw, double* __restrict__ Y, int64_t nsy, double ascale,
                           double afold, double yscale, int has_w, int groups_per_warp, int persistent, int pdist,
                           int policy, int reverse, const int* __restrict__ glist) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kq = lane & 3, nn = lane >> 2;
    const int t = kq >> 1, q = kq & 1;              // B fragment: k = (neighbour slot t, component q)
    const int r = nn >> 1, p = nn & 1;              // A / C fragment: m = (local node r, component p)
    const int slab = blockIdx.y * (NCH * 32);       // doubles per node and slab of NCH*16 columns
    const double* Xg = X + slab + (nn * 2 + q) * 2; // gather: unit (cp = 8c + nn, q)
    const int bx = reverse ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
    const int gstart = persistent ? (bx * NW + warp) : (bx * (NW * groups_per_warp) + warp);
    const int gstep = persistent ? (int)gridDim.x * NW : NW;
    const int niter = persistent ? 0x7fffffff : groups_per_warp;
    const uint64_t polX = make_policy((policy & 1) ? 1 : 0);
    const uint64_t polS = make_policy((policy & 2) ? 2 : 0);      // streams: A fragments, W
    const uint64_t polY = make_policy((policy & 4) ? 2 : ((policy & 8) ? 1 : 0));
    constexpr bool ROTC = AMODE == 1, PATTERN = AMODE == 2;
    constexpr int AW = ROTC ? 16 : 32;              // doubles per k-step of the fragment stream
    const int aoff = ROTC ? ((r * 2 + t) * 2 + (p ^ q)) : lane;
    const int colmask = (ROTC || PATTERN) ? ROTC_COLMASK : 0x7fffffff;
    const int* __restrict__ degs = reinterpret_cast<const int*>(afrag);     // PATTERN: diagonal values (row length - 1)
    double degr = 0.0;
    const int64_t eoff = slab + ((2 * kq) * 2 + p) * 2;   // C fragment: (node r, component p), column pairs 8c + 2kq (+1)

    // Y = yscale * (W + A_eff X) with A_eff = ascale * A + afold * I:  (ascale, afold, yscale) = (alpha, beta, gamma) / gamma, or
    // (1, beta / alpha, alpha) without W.  W is loaded RAW into the accumulators and the A element is post-processed (sign,
    // scale, diagonal fold) only right before its MMA, so no arithmetic sits between a load and the next loads (in-order
    // issue: the first version stalled a full L2 latency per k-step on the sign flip placed right behind the fragment load).
    auto load_a = [&](int s) -> double {
        if (PATTERN) return 0.0;
        return ldg64_stream_hint(afrag + (int64_t)s * AW + aoff, polS);
    };
    auto finish_a = [&](double v, int cw, int own) -> double {
        if (PATTERN) {
            // a * I_2 block: only p == q lanes carry a value; -1 where node r stores this neighbour, deg_r on its own column
            const bool is_own = (cw & (int)(0x80000000u | (unsigned)colmask)) == own;       // own == -1 when p != q
            const bool present = (q == p) && cw >= 0 && ((cw >> (27 + r)) & 1);
            const double a = is_own ? degr : (present ? -1.0 : 0.0);
            return fma(ascale, a, is_own ? afold : 0.0);
        }
        if (ROTC) {
            // [[a, -s b], [b, s a]]: (p,q) = (0,1) -> -s b ; (1,1) -> s a ; s = -1 when the flip bit is set
            const bool flip = (cw >> (27 + r)) & 1;
            if (q == 1 && ((p == 0) != flip)) v = -v;
        }
        return fma(ascale, v, ((cw & (int)(0x80000000u | (unsigned)colmask)) == own) ? afold : 0.0);
    };

    // glist != NULL: `ngroups` is the length of a LIST of row groups to process (row-sharded runs split the groups into those
    // that only read local rows and those that also read halo rows, so the interior launch can overlap the halo exchange)
    int gi = gstart;
    for (int it = 0; it < niter && gi < ngroups; ++it, gi += gstep) {
        const int g = glist ? __ldg(glist + gi) : gi;
        const int s0 = __ldg(kptr + g), s1 = __ldg(kptr + g + 1);
        const int own = (q == p) ? (g * 4 + r) : -1;      // column whose A entry takes the folded beta
        const int node = g * 4 + r;
        if (PATTERN) degr = (node < nbrows) ? (double)__ldg(degs + node) : 0.0;
        // L2 prefetch of what the warp's group `pdist` iterations ahead will stream from DRAM (its A fragments and W rows,
        // optionally its own X rows).  ncu showed the k-loop waiting a full DRAM latency per k-step (one k-step of
        // register prefetch, < 20 warps per SM); this turns those into L2 hits without holding registers.
        const int gpi = gi + gstep * pdist;
        const bool pf = pdist > 0 && it + pdist < niter && gpi < ngroups;
        const int gp = pf ? (glist ? __ldg(glist + gpi) : gpi) : 0;
        int ps0 = 0, ps1 = 0;
        if (pf) { ps0 = __ldg(kptr + gp); ps1 = __ldg(kptr + gp + 1); }
        double acc[NCH][2][2];
        if (has_w && node < nbrows) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const double* wp = W + (int64_t)node * nsw + eoff + 32 * c;
                const double2 w0 = ldg128_stream_hint(wp, polS), w1 = ldg128_stream_hint(wp + 4, polS);
                acc[c][0][0] = w0.x; acc[c][1][0] = w0.y; acc[c][0][1] = w1.x; acc[c][1][1] = w1.y;
            }
        } else {
#pragma unroll
            for (int c = 0; c < NCH; ++c) { acc[c][0][0] = acc[c][0][1] = acc[c][1][0] = acc[c][1][1] = 0.0; }
        }

        // register ring of KS k-steps: slot j holds step s + j; it is refilled with step s + j + KS right after its MMAs
        double a[KS];
        double2 xb[KS][NCH];
        int cc[KS], cn[KS];                           // column word of the step in slot j / of the step that will refill it
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            a[j] = 0.0; cc[j] = (int)0x80000000;
#pragma unroll
            for (int c = 0; c < NCH; ++c) xb[j][c] = make_double2(0.0, 0.0);
            if (s0 + j < s1) {
                const int cw = __ldg(kcols + (int64_t)(s0 + j) * 2 + t);
                const double* xr = Xg + (int64_t)(cw & colmask) * nsx;
#pragma unroll
                for (int c = 0; c < NCH; ++c) xb[j][c] = ldg128_hint(xr + 32 * c, polX);
                a[j] = load_a(s0 + j);
                cc[j] = cw;
            }
            cn[j] = (s0 + KS + j < s1) ? __ldg(kcols + (int64_t)(s0 + KS + j) * 2 + t) : 0;
        }
        if (pf) {
            if (lane == 0) { if (!PATTERN) prefetch_l2_bulk_hint(afrag + (int64_t)ps0 * AW, (unsigned)(ps1 - ps0) * (AW * 8u), polS); }
            else if (lane == 1) { prefetch_l2_line(kcols + (int64_t)ps0 * 2); prefetch_l2_line(kcols + (int64_t)ps1 * 2 - 1); }
            else if (lane >= 4 && lane < 12) {
                const int pn = gp * 4 + (lane & 3);
                if (pn < nbrows) {
                    if (lane < 8) { if (policy & 16) prefetch_l2_bulk_hint(X + (int64_t)pn * nsx + slab, NCH * 256u, polX); }
                    else if (has_w) prefetch_l2_bulk_hint(W + (int64_t)pn * nsw + slab, NCH * 256u, polS);
                }
            }
        }
        // The MMAs are unconditional (slots past s1 carry a = 0 and the padding bit, so they add nothing): conditional MMAs
        // made ptxas shuffle the 32 accumulator registers through ~40 MOVs per k-step.
        for (int s = s0; s < s1; s += KS) {
#pragma unroll
            for (int j = 0; j < KS; ++j) {
                const double av = finish_a(a[j], cc[j], own);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    mma_f64(acc[c][0][0], acc[c][0][1], av, xb[j][c].x);
                    mma_f64(acc[c][1][0], acc[c][1][1], av, xb[j][c].y);
                }
                const int sn = s + j + KS;
                const bool v = sn < s1;
                const int cw = cn[j];
                if (v) {
                    const double* xr = Xg + (int64_t)(cw & colmask) * nsx;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) xb[j][c] = ldg128_hint(xr + 32 * c, polX);
                    a[j] = load_a(sn);
                }
                if (!v) a[j] = 0.0;
                cc[j] = v ? cw : (int)0x80000000;
                cn[j] = (sn + KS < s1) ? __ldg(kcols + (int64_t)(sn + KS) * 2 + t) : 0;
            }
        }
        if (node < nbrows) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                double* yp = Y + (int64_t)node * nsy + eoff + 32 * c;
                stg128_hint(yp, make_double2(yscale * acc[c][0][0], yscale * acc[c][1][0]), polY);
                stg128_hint(yp + 4, make_double2(yscale * acc[c][0][1], yscale * acc[c][1][1]), polY);
            }
        }
    }
}

// PATTERN packing (AMODE 2): k-steps of 2 union neighbours straight from the R = 4 merge plan; column word = column |
// row mask << 27 (which of the group's 4 nodes store it) | bit 31 on padding; deg[i] = row length - 1 (the self loop).
__global__ void mma_pack_pattern_kernel(int ngroups, const int* __restrict__ gptr, const int2* __restrict__ uent,
                                        const int* __restrict__ kptr, int* __restrict__ kcols, int* __restrict__ bad) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    const int u0 = gptr[g], ulen = gptr[g + 1] - u0;
    const int s0 = kptr[g], ns = kptr[g + 1] - s0;
    for (int s = 0; s < ns; ++s) {
        for (int t = 0; t < 2; ++t) {
            const int u = s * 2 + t;
            const int2 ent = uent[u0 + (u < ulen ? u : ulen - 1)];
            if ((ent.x & ROTC_COLMASK) != ent.x) atomicOr(bad, 2);          // column index does not fit in 27 bits
            kcols[(int64_t)(s0 + s) * 2 + t] = (u < ulen) ? (ent.x | ((ent.y & 15) << 27)) : (ent.x | (int)0x80000000);
        }
    }
}

__global__ void row_degree_kernel(int n, const int* __restrict__ indptr, int* __restrict__ deg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) deg[i] = indptr[i + 1] - indptr[i] - 1;
}

// ROTC packing: afrag (32 doubles per k-step, fragment order) -> 16 doubles (a, b per (node r, neighbour t)) + flip bits
// in the column words.  bad (device int): set when a block is not a scaled rotation / reflection within rtol.
__global__ void mma_rotc_kernel(int64_t nk, const double* __restrict__ afrag, int* __restrict__ kcols,
                                double* __restrict__ afrag_c, int* __restrict__ bad, double rtol) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= nk) return;
    const int kq = lane & 3, nn = lane >> 2;
    const int t = kq >> 1, q = kq & 1, r = nn >> 1, p = nn & 1;
    const double v = afrag[s * 32 + lane];
    const double vq = __shfl_xor_sync(0xffffffffu, v, 1);      // same block, other column
    const double vp = __shfl_xor_sync(0xffffffffu, v, 4);      // same block, other row
    const double vpq = __shfl_xor_sync(0xffffffffu, v, 5);
    bool flip = false;
    if (p == 0 && q == 0) {                                     // m00 = v, m01 = vq, m10 = vp, m11 = vpq
        const double scale = fabs(v) + fabs(vq) + fabs(vp) + fabs(vpq);
        flip = (v * vpq - vq * vp) < 0.0;
        const double sg = flip ? -1.0 : 1.0;
        if (fabs(vpq - sg * v) > rtol * scale || fabs(vq + sg * vp) > rtol * scale) atomicOr(bad, 1);
        afrag_c[s * 16 + (r * 2 + t) * 2 + 0] = v;
        afrag_c[s * 16 + (r * 2 + t) * 2 + 1] = vp;
    }
    const unsigned fl = __ballot_sync(0xffffffffu, flip);       // bit (8r + 2t) = flip of block (r, t)
    if (lane < 2) {                                             // lane = t
        int w = kcols[s * 2 + lane];
        if ((w & ROTC_COLMASK) != (w & 0x7fffffff)) atomicOr(bad, 2);      // column index does not fit in 27 bits
        for (int rr = 0; rr < 4; ++rr) if ((fl >> (8 * rr + 2 * lane)) & 1u) w |= 1 << (27 + rr);
        kcols[s * 2 + lane] = w;
    }
}

// Row-major panel (rows 2*node + q, leading dimension ldv) <-> node-contiguous panel (node stride ns).  One thread per
// 16-byte unit (column pair cp, component q).
template <bool TO_NATIVE>
__global__ void native_convert_kernel(int64_t nbrows, int ncp, double* __restrict__ V, int64_t ldv, double* __restrict__ Xn,
                                      int64_t ns) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per_node = 2 * (int64_t)ncp;
    if (idx >= nbrows * per_node) return;
    const int64_t node = idx / per_node;
    const int rem = (int)(idx - node * per_node);
    const int q = rem / ncp, cp = rem - q * ncp;          // consecutive threads walk along a V row
    double2* v = reinterpret_cast<double2*>(V + (2 * node + q) * ldv) + cp;
    double2* x = reinterpret_cast<double2*>(Xn + node * ns) + (cp * 2 + q);
    if (TO_NATIVE) *x = *v; else *v = *x;
}

int native_convert(Handle* h, bool to_native, int nbrows, int ncols, double* V, int64_t ldv, double* Xn, int64_t ns) {
    RVGP_REQUIRE(h, ncols % 2 == 0 && ldv % 2 == 0 && ns % 2 == 0 && ns >= 2 * ncols && (uintptr_t)V % 16 == 0 &&
                        (uintptr_t)Xn % 16 == 0, "native_convert: even ncols / strides and 16-byte aligned buffers");
    if (nbrows == 0) return RVGP_OK;
    const int64_t total = (int64_t)nbrows * ncols;
    if (to_native) native_convert_kernel<true><<<cdiv(total, 256), 256, 0, h->stream>>>(nbrows, ncols / 2, V, ldv, Xn, ns);
    else native_convert_kernel<false><<<cdiv(total, 256), 256, 0, h->stream>>>(nbrows, ncols / 2, V, ldv, Xn, ns);
    RVGP_LAUNCH_OK(h, "native_convert_kernel");
    return RVGP_OK;
}

int spmm_mma_native_dispatch_list(Handle* h, int nbrows, const int* kptr, const int* kcols, const double* afrag, int rotc,
                                  const double* X, int64_t nsx, const double* W, int64_t nsw, double* Y, int64_t nsy,
                                  int ncols, double alpha, double beta, double gamma, int reverse, const int* glist, int nlist);

int spmm_mma_native_dispatch(Handle* h, int nbrows, const int* kptr, const int* kcols, const double* afrag, int rotc,
                             const double* X, int64_t nsx, const double* W, int64_t nsw, double* Y, int64_t nsy,
                             int ncols, double alpha, double beta, double gamma, int reverse) {
    return spmm_mma_native_dispatch_list(h, nbrows, kptr, kcols, afrag, rotc, X, nsx, W, nsw, Y, nsy, ncols, alpha, beta, gamma,
                                         reverse, nullptr, 0);
}

// glist (nullable, device int32[nlist]): process only these row groups (groups of 4 block rows), in list order.
int spmm_mma_native_dispatch_list(Handle* h, int nbrows, const int* kptr, const int* kcols, const double* afrag, int rotc,
                                  const double* X, int64_t nsx, const double* W, int64_t nsw, double* Y, int64_t nsy,
                                  int ncols, double alpha, double beta, double gamma, int reverse, const int* glist, int nlist) {
    RVGP_REQUIRE(h, ncols >= 16 && ncols % 16 == 0, "spmm_mma_native: ncols must be a multiple of 16");
    RVGP_REQUIRE(h, nsx % 2 == 0 && nsy % 2 == 0 && (W == nullptr || nsw % 2 == 0) && (uintptr_t)X % 16 == 0 &&
                        (uintptr_t)Y % 16 == 0 && (uintptr_t)W % 16 == 0,
                 "spmm_mma_native: buffers must be 16-byte aligned with even node strides");
    RVGP_REQUIRE(h, gamma == 0.0 || W != nullptr, "spmm_mma_native: W required when gamma != 0");
    RVGP_REQUIRE(h, Y != X && Y != W, "spmm_mma_native: Y must not alias X or W");
    RVGP_REQUIRE(h, alpha != 0.0, "spmm_mma_native: alpha must be non-zero (beta is folded into the matrix as beta / alpha)");
    if (nbrows == 0 || (glist != nullptr && nlist == 0)) return RVGP_OK;
    const int ngroups = glist ? nlist : cdiv(nbrows, 4);
    const int gpw = h->mma_gpw;                       // 0 = persistent warp-strided schedule
    const int var = h->mma_variant;
    const int nch = (ncols % 64 == 0) ? 4 : ((ncols % 32 == 0) ? 2 : 1);
    const int nslab = ncols / (16 * nch);
    const int has_w = gamma != 0.0;
    const double ascale = has_w ? alpha / gamma : 1.0, afold = has_w ? beta / gamma : beta / alpha, yscale = has_w ? gamma : alpha;
#define RVGP_MMAN(NCH, KS, NT, MINB)                                                                                   \
    do {                                                                                                               \
        const int nw_ = NT / 32;                                                                                       \
        const int gx = gpw > 0 ? cdiv(ngroups, nw_ * gpw) : (int)std::min<long long>(cdiv(ngroups, nw_), std::max(1, h->sm_count * MINB / nslab)); \
        dim3 grid(gx, nslab);                                                                                          \
        if (rotc == 2) bsr_spmm_mma_native_kernel<NCH, KS, NT, MINB, 2><<<grid, NT, 0, h->stream>>>(                   \
            ngroups, nbrows, kptr, kcols, afrag, X, nsx, W, nsw, Y, nsy, ascale, afold, yscale, has_w, gpw, gpw <= 0,           \
            h->mma_prefetch, h->mma_stream_policy, reverse, glist);                                                    \
        else if (rotc) bsr_spmm_mma_native_kernel<NCH, KS, NT, MINB, 1><<<grid, NT, 0, h->stream>>>(                   \
            ngroups, nbrows, kptr, kcols, afrag, X, nsx, W, nsw, Y, nsy, ascale, afold, yscale, has_w, gpw, gpw <= 0,           \
            h->mma_prefetch, h->mma_stream_policy, reverse, glist);                                                    \
        else bsr_spmm_mma_native_kernel<NCH, KS, NT, MINB, 0><<<grid, NT, 0, h->stream>>>(                             \
            ngroups, nbrows, kptr, kcols, afrag, X, nsx, W, nsw, Y, nsy, ascale, afold, yscale, has_w, gpw, gpw <= 0,           \
            h->mma_prefetch, h->mma_stream_policy, reverse, glist);                                                    \
    } while (0)
    // (chunks of 16 columns, register-ring depth, threads per CTA, CTAs per SM); measured at C4 size, persistent schedule,
    // policy 7 (profiles/r01_spmm_mma_sweep.txt): 1: 0.828 ms | 0: 0.854 | 6: 0.851 | 5: 0.858 | 3: 0.864 | 2: 0.882
    if (nch == 4) {
        if (var == 0) RVGP_MMAN(4, 2, 128, 4); else if (var == 2) RVGP_MMAN(4, 3, 128, 4); else if (var == 3) RVGP_MMAN(4, 3, 256, 2);
        else if (var == 5) RVGP_MMAN(4, 4, 128, 3); else if (var == 6) RVGP_MMAN(4, 3, 384, 1); else RVGP_MMAN(4, 2, 256, 2);
    } else if (nch == 2) {
        // 32 native columns = the 64-column row-major panels of the scalar Laplacian (AMODE 2) in the Krylov solver.
        // (chunks, ring depth, threads, CTAs/SM) swept with tools/profile_k9.py, profiles/r02_k9_sweep.txt
        const int v2 = h->mma_variant_n2;
        if (v2 == 1) RVGP_MMAN(2, 2, 256, 3); else if (v2 == 2) RVGP_MMAN(2, 3, 128, 6); else if (v2 == 3) RVGP_MMAN(2, 4, 128, 4);
        else if (v2 == 4) RVGP_MMAN(2, 3, 256, 3); else if (v2 == 5) RVGP_MMAN(2, 4, 256, 2); else if (v2 == 6) RVGP_MMAN(2, 4, 128, 6);
        else RVGP_MMAN(2, 2, 128, 6);
    } else {
        RVGP_MMAN(1, 2, 128, 6);
    }
#undef RVGP_MMAN
    RVGP_LAUNCH_OK(h, "bsr_spmm_mma_native_kernel");
    return RVGP_OK;
}

}  // namespace rvgp

using namespace rvgp;

// Pack the k-step plan of the MMA SpMM from a row-group merge plan (rvgp_bsr_merge_plan with R = 8/d).
// kptr (ngroups+1): exclusive prefix of ceil(ulen_g / (4/d)); kcols: kptr[ngroups] * (4/d) int32; afrag: kptr[ngroups] * 32
// doubles.  vals == NULL (d == 1): unit-weight graph Laplacian pattern.
extern "C" int rvgp_bsr_mma_pack(rvgp_handle_t hh, int nbrows, int d, const int32_t* indptr, const int32_t* indices,
                                 const double* vals, const int32_t* gptr, const int32_t* uent, const int32_t* kptr,
                                 int32_t* kcols, double* afrag) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, d == 1 || d == 2, "mma_pack: block size d must be 1 or 2");
    RVGP_REQUIRE(h, vals != nullptr || d == 1, "mma_pack: pattern mode (vals == NULL) needs d == 1");
    const int ngroups = cdiv(nbrows, 8 / d);
    if (ngroups == 0) return RVGP_OK;
    const int2* ue = reinterpret_cast<const int2*>(uent);
    if (d == 2) mma_pack_kernel<2><<<cdiv(ngroups, 4), 128, 0, h->stream>>>(nbrows, ngroups, indptr, indices, vals, gptr, ue, kptr, kcols, afrag);
    else        mma_pack_kernel<1><<<cdiv(ngroups, 4), 128, 0, h->stream>>>(nbrows, ngroups, indptr, indices, vals, gptr, ue, kptr, kcols, afrag);
    RVGP_LAUNCH_OK(h, "mma_pack_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_bsr_spmm_mma_f64(rvgp_handle_t hh, int nbrows, int d, const int32_t* kptr, const int32_t* kcols,
                                     const double* afrag, const double* X, int64_t ldx, const double* W, int64_t ldw,
                                     double* Y, int64_t ldy, int ncols, double alpha, double beta, double gamma) {
    return spmm_mma_dispatch(H(hh), nbrows, d, kptr, kcols, afrag, X, ldx, W, ldw, Y, ldy, ncols, alpha, beta, gamma);
}

// Same product on NODE-CONTIGUOUS panels (d == 2 only): element (2*node + q, 2*cp + e) of the block vector lives at
// Xn[node * ns + (cp * 2 + q) * 2 + e]; ns* are the node strides in doubles (>= 2 * ncols).  beta is folded into the
// diagonal of A as beta / alpha, so alpha must be non-zero.  rotc != 0: afrag / kcols are in the compact form written
// by rvgp_bsr_mma_rotc.  reverse != 0 walks the row groups from the last to the first (alternate it between the steps
// of a recurrence so that each launch starts on the rows the previous one left in L2).
extern "C" int rvgp_bsr_spmm_mma_native_f64(rvgp_handle_t hh, int nbrows, const int32_t* kptr, const int32_t* kcols,
                                            const double* afrag, int rotc, const double* Xn, int64_t nsx,
                                            const double* Wn, int64_t nsw, double* Yn, int64_t nsy, int ncols,
                                            double alpha, double beta, double gamma, int reverse) {
    return spmm_mma_native_dispatch(H(hh), nbrows, kptr, kcols, afrag, rotc, Xn, nsx, Wn, nsw, Yn, nsy, ncols, alpha, beta,
                                    gamma, reverse);
}

// Compact the k-step plan of a d == 2 matrix whose blocks are all scaled rotations / reflections: afrag (nk * 32) ->
// afrag_c (nk * 16), flip bits into kcols (in place, bits 27..30).  bad_flag (device int32, zeroed here): bit0 = some block
// is not of that form within rtol, bit1 = a column index needs more than 27 bits; when non-zero keep the full plan
// (kcols must then be re-packed, it has been modified).
extern "C" int rvgp_bsr_mma_rotc(rvgp_handle_t hh, int64_t nk, const double* afrag, int32_t* kcols, double* afrag_c,
                                 int32_t* bad_flag, double rtol) {
    Handle* h = H(hh);
    RVGP_CUDA_OK(h, cudaMemsetAsync(bad_flag, 0, sizeof(int), h->stream));
    if (nk == 0) return RVGP_OK;
    mma_rotc_kernel<<<cdiv(nk, 4), 128, 0, h->stream>>>(nk, afrag, kcols, afrag_c, bad_flag, rtol);
    RVGP_LAUNCH_OK(h, "mma_rotc_kernel");
    return RVGP_OK;
}

// Layout conversion between a row-major (2*nbrows x ncols, leading dimension ldv) panel and the node-contiguous panel
// of rvgp_bsr_spmm_mma_native_f64 (node stride ns >= 2*ncols).  to_native != 0: V -> Xn, else Xn -> V.
extern "C" int rvgp_panel_native_f64(rvgp_handle_t hh, int to_native, int nbrows, int ncols, double* V, int64_t ldv,
                                     double* Xn, int64_t ns) {
    return native_convert(H(hh), to_native != 0, nbrows, ncols, V, ldv, Xn, ns);
}

// PATTERN plan of the scalar unit-weight Laplacian for the native kernel (rotc == 2 in rvgp_bsr_spmm_mma_native_f64 /
// rvgp_cheb_filter_mma_f64): gptr / uent from rvgp_bsr_merge_plan with R = 4; kptr (ngroups + 1) = exclusive prefix of
// ceil(ulen_g / 2); kcols: 2 * kptr[ngroups] int32; deg: nbrows int32 (pass it as `afrag`).  bad_flag (device int32, zeroed
// here): bit1 = a column index needs more than 27 bits.
extern "C" int rvgp_bsr_mma_pack_pattern(rvgp_handle_t hh, int nbrows, const int32_t* indptr, const int32_t* gptr,
                                         const int32_t* uent, const int32_t* kptr, int32_t* kcols, int32_t* deg,
                                         int32_t* bad_flag) {
    Handle* h = H(hh);
    RVGP_CUDA_OK(h, cudaMemsetAsync(bad_flag, 0, sizeof(int), h->stream));
    const int ngroups = cdiv(nbrows, 4);
    if (ngroups == 0) return RVGP_OK;
    mma_pack_pattern_kernel<<<cdiv(ngroups, 128), 128, 0, h->stream>>>(ngroups, gptr, reinterpret_cast<const int2*>(uent), kptr,
                                                                        kcols, bad_flag);
    RVGP_LAUNCH_OK(h, "mma_pack_pattern_kernel");
    row_degree_kernel<<<cdiv(nbrows, 256), 256, 0, h->stream>>>(nbrows, indptr, deg);
    RVGP_LAUNCH_OK(h, "row_degree_kernel");
    return RVGP_OK;
}

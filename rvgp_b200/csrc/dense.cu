// K10 / K13: FP64 dense kernels.
//
//  * rvgp_dgemm_f64: register-tiled DGEMM (128x128x16 CTA tile, 8x8 per thread, DFMA pipe) with
//    selectable operand layouts, optional K-scaling (spectral density S, reference kernels.py:61) and a
//    deterministic split-K for tall-skinny Gram products V^T W (orthogonalisation / Rayleigh-Ritz of the
//    block eigensolver that replaces ARPACK's reorthogonalisation, reference geometry.py:73).
//    tcgen05 has no FP64 kind, so this is the FP64-FMA-pipe roofline (DESIGN.md "K10/K13").
//  * column-wise reductions and small utilities over tall block vectors (HBM-bound, read once).
#include "common.cuh"

namespace rvgp {

constexpr int BM = 128, BN = 128, BK = 16, PADM = 2;

template <bool A_KMAJOR, bool B_KMAJOR, bool HAS_SCALE>
__global__ void __launch_bounds__(256)
dgemm_kernel(int M, int N, int64_t K, double alpha, const double* __restrict__ A, int64_t lda,
             const double* __restrict__ B, int64_t ldb, const double* __restrict__ scale_k,
             double* __restrict__ C, int64_t ldc, int split_k, double* __restrict__ ws, double beta, int lower_only) {
    if (lower_only && blockIdx.x * BN > blockIdx.y * BM + (BM - 1)) return;   // SYRK: tiles strictly above the diagonal
    __shared__ __align__(16) double As[BK][BM + PADM];
    __shared__ __align__(16) double Bs[BK][BN + PADM];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // K range of this split (multiple of BK)
    const int64_t ktiles = (K + BK - 1) / BK;
    const int64_t tiles_per = (ktiles + split_k - 1) / split_k;
    const int64_t kt0 = (int64_t)blockIdx.z * tiles_per;
    const int64_t kt1 = (kt0 + tiles_per < ktiles) ? kt0 + tiles_per : ktiles;

    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

    double ra[8], rb[8], rs[HAS_SCALE ? 8 : 1];
    auto load_tiles = [&](int64_t kt) {
        const int64_t k0 = kt * BK;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int i, k;
            if (A_KMAJOR) { k = t & 15; i = (t >> 4) + 16 * r; }
            else          { i = t & 127; k = (t >> 7) + 2 * r; }
            const int64_t gk = k0 + k;
            const int gi = m0 + i;
            const bool ok = (gi < M) && (gk < K);
            // branch-free predicated loads: all eight requests of a thread are in flight together (the first version
            // branched per element and serialised them -- ncu: long-scoreboard stalls dominated)
            const double* pa = A_KMAJOR ? (A + (int64_t)(ok ? gi : 0) * lda + (ok ? gk : 0)) : (A + (ok ? gk : 0) * lda + (ok ? gi : 0));
            ra[r] = ok ? __ldg(pa) : 0.0;
            if (HAS_SCALE) rs[r] = ok ? __ldg(scale_k + gk) : 0.0;
        }
        if (HAS_SCALE) {
#pragma unroll
            for (int r = 0; r < 8; ++r) ra[r] *= rs[r];
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int j, k;
            if (B_KMAJOR) { k = t & 15; j = (t >> 4) + 16 * r; }
            else          { j = t & 127; k = (t >> 7) + 2 * r; }
            const int64_t gk = k0 + k;
            const int gj = n0 + j;
            const bool ok = (gj < N) && (gk < K);
            const double* pb = B_KMAJOR ? (B + (int64_t)(ok ? gj : 0) * ldb + (ok ? gk : 0)) : (B + (ok ? gk : 0) * ldb + (ok ? gj : 0));
            rb[r] = ok ? __ldg(pb) : 0.0;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int i, k;
            if (A_KMAJOR) { k = t & 15; i = (t >> 4) + 16 * r; }
            else          { i = t & 127; k = (t >> 7) + 2 * r; }
            As[k][i] = ra[r];
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int j, k;
            if (B_KMAJOR) { k = t & 15; j = (t >> 4) + 16 * r; }
            else          { j = t & 127; k = (t >> 7) + 2 * r; }
            Bs[k][j] = rb[r];
        }
    };

    if (kt0 < kt1) load_tiles(kt0);
    for (int64_t kt = kt0; kt < kt1; ++kt) {
        store_tiles();
        __syncthreads();
        if (kt + 1 < kt1) load_tiles(kt + 1);   // global loads in flight while the tile is consumed
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            double a[8], b[8];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const double2 v = *reinterpret_cast<const double2*>(&As[kk][2 * ty + 32 * ii]);
                a[2 * ii] = v.x; a[2 * ii + 1] = v.y;
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const double2 v = *reinterpret_cast<const double2*>(&Bs[kk][2 * tx + 32 * jj]);
                b[2 * jj] = v.x; b[2 * jj + 1] = v.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    double* out = C;
    int64_t ldo = ldc;
    double sc = alpha;
    if (split_k > 1) { out = ws + (int64_t)blockIdx.z * M * N; ldo = N; sc = 1.0; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gi = m0 + 2 * ty + 32 * (i >> 1) + (i & 1);
        if (gi >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int gj = n0 + 2 * tx + 32 * (j >> 1) + (j & 1);
            if (gj < N) {
                double v = sc * acc[i][j];
                if (beta != 0.0 && split_k == 1) v = fma(beta, out[(int64_t)gi * ldo + gj], v);
                out[(int64_t)gi * ldo + gj] = v;
            }
        }
    }
}

// ---- DMMA variant: same tiling / loaders, inner product on the FP64 tensor path (mma.sync.m8n8k4.f64) ---------
// CTA tile 128 x 64 x 16, 8 warps as 4 (m) x 2 (n), warp tile 32 x 32 = 4 x 4 MMA tiles.  Fragment layout (PTX ISA):
//   A (8x4, row): a0 -> row = lane>>2, k = lane&3      B (4x8, col): b0 -> k = lane&3, n = lane>>2
//   C/D (8x8):    c0,c1 -> row = lane>>2, cols = 2*(lane&3) + {0,1}
// (tcgen05 has no FP64 kind; this legacy warp-level MMA is the only FP64 tensor path on sm_100a.)
constexpr int DBN = 64, DPAD = 4;   // row stride = 4 (mod 16) doubles: the (k = lane&3, m = lane>>2) fragment loads are bank-conflict free

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <bool A_KMAJOR, bool B_KMAJOR, bool HAS_SCALE>
__global__ void __launch_bounds__(256, 2)
dgemm_dmma_kernel(int M, int N, int64_t K, double alpha, const double* __restrict__ A, int64_t lda,
                  const double* __restrict__ B, int64_t ldb, const double* __restrict__ scale_k,
                  double* __restrict__ C, int64_t ldc, int split_k, double* __restrict__ ws, double beta, int lower_only) {
    if (lower_only && blockIdx.x * DBN > blockIdx.y * BM + (BM - 1)) return;
    __shared__ __align__(16) double As[BK][BM + DPAD];
    __shared__ __align__(16) double Bs[BK][DBN + DPAD];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int wm = warp >> 1, wn = warp & 1;                 // 4 x 2 warps
    const int gid = lane >> 2, tig = lane & 3;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * DBN;
    const int64_t ktiles = (K + BK - 1) / BK;
    const int64_t tiles_per = (ktiles + split_k - 1) / split_k;
    const int64_t kt0 = (int64_t)blockIdx.z * tiles_per;
    const int64_t kt1 = (kt0 + tiles_per < ktiles) ? kt0 + tiles_per : ktiles;

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    double ra[8], rb[4], rs[HAS_SCALE ? 8 : 1];
    auto load_tiles = [&](int64_t kt) {
        const int64_t k0 = kt * BK;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int i, k;
            if (A_KMAJOR) { k = t & 15; i = (t >> 4) + 16 * r; }
            else          { i = t & 127; k = (t >> 7) + 2 * r; }
            const int64_t gk = k0 + k;
            const int gi = m0 + i;
            const bool ok = (gi < M) && (gk < K);
            // branch-free predicated loads: all eight requests of a thread are in flight together (the first version
            // branched per element and serialised them -- ncu: long-scoreboard stalls dominated)
            const double* pa = A_KMAJOR ? (A + (int64_t)(ok ? gi : 0) * lda + (ok ? gk : 0)) : (A + (ok ? gk : 0) * lda + (ok ? gi : 0));
            ra[r] = ok ? __ldg(pa) : 0.0;
            if (HAS_SCALE) rs[r] = ok ? __ldg(scale_k + gk) : 0.0;
        }
        if (HAS_SCALE) {
#pragma unroll
            for (int r = 0; r < 8; ++r) ra[r] *= rs[r];
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int j, k;
            if (B_KMAJOR) { k = t & 15; j = (t >> 4) + 16 * r; }
            else          { j = t & 63; k = (t >> 6) + 4 * r; }
            const int64_t gk = k0 + k;
            const int gj = n0 + j;
            const bool ok = (gj < N) && (gk < K);
            const double* pb = B_KMAJOR ? (B + (int64_t)(ok ? gj : 0) * ldb + (ok ? gk : 0)) : (B + (ok ? gk : 0) * ldb + (ok ? gj : 0));
            rb[r] = ok ? __ldg(pb) : 0.0;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int i, k;
            if (A_KMAJOR) { k = t & 15; i = (t >> 4) + 16 * r; }
            else          { i = t & 127; k = (t >> 7) + 2 * r; }
            As[k][i] = ra[r];
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int j, k;
            if (B_KMAJOR) { k = t & 15; j = (t >> 4) + 16 * r; }
            else          { j = t & 63; k = (t >> 6) + 4 * r; }
            Bs[k][j] = rb[r];
        }
    };

    if (kt0 < kt1) load_tiles(kt0);
    for (int64_t kt = kt0; kt < kt1; ++kt) {
        store_tiles();
        __syncthreads();
        if (kt + 1 < kt1) load_tiles(kt + 1);
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k4 + tig][wm * 32 + i * 8 + gid];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k4 + tig][wn * 32 + j * 8 + gid];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }

    double* out = C;
    int64_t ldo = ldc;
    double sc = alpha;
    if (split_k > 1) { out = ws + (int64_t)blockIdx.z * M * N; ldo = N; sc = 1.0; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = m0 + wm * 32 + i * 8 + gid;
        if (gi >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int gj = n0 + wn * 32 + j * 8 + 2 * tig + c;
                if (gj < N) {
                    double v = sc * acc[i][j][c];
                    if (beta != 0.0 && split_k == 1) v = fma(beta, out[(int64_t)gi * ldo + gj], v);
                    out[(int64_t)gi * ldo + gj] = v;
                }
            }
    }
}

// ---- pipelined DMMA variant (round 2): same CTA / warp tiling as dgemm_dmma_kernel, operands streamed global -> shared by
// cp.async (16 B) through a 3-stage ring, one __syncthreads per k-tile.  The register-staged kernel above sits at ~17 TFLOP/s
// on the eigensolver's two skinny shapes with long-scoreboard as its first stall (profiles/r01_dgemm_dmma_summary.txt: every
// warp waits for its own global loads before it can store the tile); here the loads of tile kt+2 are in flight while tile kt
// is consumed and no warp ever waits on a register.  Layouts: B is N-major (k rows of n-contiguous doubles, the only layout
// the eigensolver uses); A is M-major (Gram V^T W: k rows of m-contiguous doubles) or K-major (update W -= V C: m rows of
// k-contiguous doubles, kept in that orientation in shared memory).  Row strides = 4 (mod 16) doubles keep the
// (gid, tig) fragment reads conflict-free in both orientations.  Needs 16-byte aligned operands and even extents (dispatch).
constexpr int PST = 3;
constexpr int PA_M = BK * (BM + DPAD);
constexpr int PA_K = BM * (BK + DPAD);
constexpr int PA_SZ = PA_K > PA_M ? PA_K : PA_M;
constexpr int PB_SZ = BK * (DBN + DPAD);
constexpr int PSTAGE = PA_SZ + PB_SZ;
constexpr size_t PIPE_SMEM = (size_t)PST * PSTAGE * sizeof(double);

__device__ __forceinline__ void cp_async16(double* smem, const double* gmem, bool ok) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = ok ? 16 : 0;                       // src-size 0: the 16 bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(s), "l"(gmem), "r"(sz) : "memory");
}

template <bool A_KMAJOR>
__global__ void __launch_bounds__(256, 2)
dgemm_dmma_pipe_kernel(int M, int N, int64_t K, double alpha, const double* __restrict__ A, int64_t lda,
                       const double* __restrict__ B, int64_t ldb, double* __restrict__ C, int64_t ldc, int split_k,
                       double* __restrict__ ws, double beta, int lower_only) {
    if (lower_only && blockIdx.x * DBN > blockIdx.y * BM + (BM - 1)) return;
    extern __shared__ __align__(16) double psm[];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int gid = lane >> 2, tig = lane & 3;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * DBN;
    const int64_t ktiles = (K + BK - 1) / BK;
    const int64_t tiles_per = (ktiles + split_k - 1) / split_k;
    const int64_t kt0 = (int64_t)blockIdx.z * tiles_per;
    const int64_t kt1 = (kt0 + tiles_per < ktiles) ? kt0 + tiles_per : ktiles;
    const int nk = kt1 > kt0 ? (int)(kt1 - kt0) : 0;

    auto issue = [&](int64_t kt, int st) {
        double* As = psm + st * PSTAGE;
        double* Bs = As + PA_SZ;
        const int64_t k0 = kt * BK;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int c = t + 256 * r;
            if (A_KMAJOR) {
                const int m = c >> 3, kp = (c & 7) * 2;
                const int gi = m0 + m;
                const int64_t gk = k0 + kp;
                const bool ok = (gi < M) && (gk < K);
                cp_async16(As + m * (BK + DPAD) + kp, ok ? A + (int64_t)gi * lda + gk : A, ok);
            } else {
                const int k = c >> 6, mp = (c & 63) * 2;
                const int gi = m0 + mp;
                const int64_t gk = k0 + k;
                const bool ok = (gi < M) && (gk < K);
                cp_async16(As + k * (BM + DPAD) + mp, ok ? A + gk * lda + gi : A, ok);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int c = t + 256 * r;
            const int k = c >> 5, np = (c & 31) * 2;
            const int gj = n0 + np;
            const int64_t gk = k0 + k;
            const bool ok = (gj < N) && (gk < K);
            cp_async16(Bs + k * (DBN + DPAD) + np, ok ? B + gk * ldb + gj : B, ok);
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

#pragma unroll
    for (int s = 0; s < PST - 1; ++s) {
        if (s < nk) issue(kt0 + s, s);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int it = 0; it < nk; ++it) {
        asm volatile("cp.async.wait_group %0;" :: "n"(PST - 2) : "memory");
        __syncthreads();                               // tile `it` visible to all; everyone is done with tile it-1's stage
        if (it + PST - 1 < nk) issue(kt0 + it + PST - 1, (it + PST - 1) % PST);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double* As = psm + (it % PST) * PSTAGE;
        const double* Bs = As + PA_SZ;
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                a[i] = A_KMAJOR ? As[(wm * 32 + i * 8 + gid) * (BK + DPAD) + k4 + tig] : As[(k4 + tig) * (BM + DPAD) + wm * 32 + i * 8 + gid];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(k4 + tig) * (DBN + DPAD) + wn * 32 + j * 8 + gid];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");

    double* out = C;
    int64_t ldo = ldc;
    double sc = alpha;
    if (split_k > 1) { out = ws + (int64_t)blockIdx.z * M * N; ldo = N; sc = 1.0; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = m0 + wm * 32 + i * 8 + gid;
        if (gi >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gj = n0 + wn * 32 + j * 8 + 2 * tig;
            if (gj + 1 < N) {                          // N is even (dispatch): the pair is in range together
                double2 v = make_double2(sc * acc[i][j][0], sc * acc[i][j][1]);
                double2* po = reinterpret_cast<double2*>(out + (int64_t)gi * ldo + gj);
                if (beta != 0.0 && split_k == 1) { const double2 o = *po; v.x = fma(beta, o.x, v.x); v.y = fma(beta, o.y, v.y); }
                *po = v;
            }
        }
    }
}

// ---- small-shape DMMA variant: 32 x 32 x 32 CTA tile, 4 warps (2 x 2), warp tile 16 x 16 --------------------------------------
// The k x k (k ~ 500) Cholesky panels and solves of the rank-k GP evaluation are a few hundred rows by 64..256 columns: with the
// 128 x 64 tile they occupy 4-12 of the 148 SMs and each CTA then needs >= 17 us for a K = 256 product (one SM's FP64 rate).
// Sixteen times more, smaller CTAs put the same product on 16 x as many SMs; operands stay in the orientation they have in
// memory (row stride 36 = 4 mod 16 doubles: conflict-free stores and (gid, tig) fragment reads in both orientations).
constexpr int SBM = 32, SBN = 32, SBK = 32, SLD = SBK + 4;

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(128)
dgemm_small_kernel(int M, int N, int64_t K, double alpha, const double* __restrict__ A, int64_t lda,
                   const double* __restrict__ B, int64_t ldb, double* __restrict__ C, int64_t ldc, double beta, int lower_only) {
    if (lower_only && blockIdx.x * SBN > blockIdx.y * SBM + (SBM - 1)) return;
    __shared__ double As[SBK * SLD];      // K-major source: [m][k], otherwise [k][m]
    __shared__ double Bs[SBK * SLD];      // K-major source: [n][k], otherwise [k][n]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int gid = lane >> 2, tig = lane & 3;
    const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
    const int64_t ktiles = (K + SBK - 1) / SBK;
    double acc[2][2][2] = {};
    double ra[8], rb[8];
    // element (fast index f = lane, slow index s = warp + 4 r): for a K-major operand f runs along k, otherwise along m / n
    auto load_tiles = [&](int64_t kt) {
        const int64_t k0 = kt * SBK;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int f = lane, sl = warp + 4 * r;
            {
                const int64_t gk = k0 + (A_KMAJOR ? f : sl);
                const int gi = m0 + (A_KMAJOR ? sl : f);
                const bool ok = (gi < M) && (gk < K);
                const double* pa = A_KMAJOR ? (A + (int64_t)(ok ? gi : 0) * lda + (ok ? gk : 0)) : (A + (ok ? gk : 0) * lda + (ok ? gi : 0));
                ra[r] = ok ? __ldg(pa) : 0.0;
            }
            {
                const int64_t gk = k0 + (B_KMAJOR ? f : sl);
                const int gj = n0 + (B_KMAJOR ? sl : f);
                const bool ok = (gj < N) && (gk < K);
                const double* pb = B_KMAJOR ? (B + (int64_t)(ok ? gj : 0) * ldb + (ok ? gk : 0)) : (B + (ok ? gk : 0) * ldb + (ok ? gj : 0));
                rb[r] = ok ? __ldg(pb) : 0.0;
            }
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            As[(warp + 4 * r) * SLD + lane] = ra[r];
            Bs[(warp + 4 * r) * SLD + lane] = rb[r];
        }
    };
    if (ktiles > 0) load_tiles(0);
    for (int64_t kt = 0; kt < ktiles; ++kt) {
        store_tiles();
        __syncthreads();
        if (kt + 1 < ktiles) load_tiles(kt + 1);
#pragma unroll
        for (int k4 = 0; k4 < SBK; k4 += 4) {
            double a[2], b[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int m = wm * 16 + i * 8 + gid;
                a[i] = A_KMAJOR ? As[m * SLD + k4 + tig] : As[(k4 + tig) * SLD + m];
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = wn * 16 + j * 8 + gid;
                b[j] = B_KMAJOR ? Bs[n * SLD + k4 + tig] : Bs[(k4 + tig) * SLD + n];
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int gi = m0 + wm * 16 + i * 8 + gid;
        if (gi >= M) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int gj = n0 + wn * 16 + j * 8 + 2 * tig + c;
                if (gj < N) {
                    double v = alpha * acc[i][j][c];
                    if (beta != 0.0) v = fma(beta, C[(int64_t)gi * ldc + gj], v);
                    C[(int64_t)gi * ldc + gj] = v;
                }
            }
    }
}

__global__ void splitk_reduce_kernel(int M, int N, int split_k, double alpha, const double* __restrict__ ws,
                                     double* __restrict__ C, int64_t ldc, double beta) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * N) return;
    double s = 0.0;
    for (int z = 0; z < split_k; ++z) s += ws[(int64_t)z * M * N + idx];   // fixed order: deterministic
    double* c = C + (idx / N) * ldc + (idx % N);
    *c = (beta != 0.0) ? fma(beta, *c, alpha * s) : alpha * s;
}

// ---- column reductions ------------------------------------------------------------------------------
// stage 1: each CTA reduces a contiguous slab of rows for all columns; stage 2: fixed-order sum of slabs.
constexpr int RED_ROWS_PER_CTA = 2048;

template <int MODE>  // 0: sum A*B   1: sum (W - theta*V)^2  (A=W, B=V)
__global__ void __launch_bounds__(256)
colreduce_stage1(int64_t nrows, int ncols, const double* __restrict__ A, int64_t lda,
                 const double* __restrict__ B, int64_t ldb, const double* __restrict__ theta,
                 double* __restrict__ partial) {
    // thread layout: 32 columns x 8 row-lanes per pass over a 32-column panel
    __shared__ double sm[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * RED_ROWS_PER_CTA;
    const int64_t r1 = (r0 + RED_ROWS_PER_CTA < nrows) ? r0 + RED_ROWS_PER_CTA : nrows;
    // one CTA per (row slab, 32-column panel): round 1 looped over the panels inside the CTA, which left a C2-sized reduction
    // (35 k rows = 17 slabs) on 17 CTAs and made this kernel 14-20 % of the C2 kernel time (profiles/r02_launch_shares_c2.txt)
    {
        const int c0 = blockIdx.y * 32;
        const int c = c0 + cx;
        double s = 0.0;
        if (c < ncols) {
            const double th = (MODE == 1) ? __ldg(theta + c) : 0.0;
            for (int64_t r = r0 + ry; r < r1; r += 8) {
                const double a = __ldg(A + r * lda + c), b = B ? __ldg(B + r * ldb + c) : 1.0;
                if (MODE == 0) s = fma(a, b, s);
                else { const double d = a - th * b; s = fma(d, d, s); }
            }
        }
        sm[ry][cx] = s;
        __syncthreads();
        if (ry == 0 && c < ncols) {
            double tot = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) tot += sm[k][cx];
            partial[(int64_t)blockIdx.x * ncols + c] = tot;
        }
        __syncthreads();
    }
}

__global__ void colreduce_stage2(int nslabs, int ncols, const double* __restrict__ partial, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    double s = 0.0;
    for (int k = 0; k < nslabs; ++k) s += partial[(int64_t)k * ncols + c];
    out[c] = s;
}

__global__ void colscale_kernel(int64_t nrows, int ncols, double* __restrict__ A, int64_t lda, const double* __restrict__ s) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrows * ncols) return;
    const int64_t r = idx / ncols;
    const int c = (int)(idx % ncols);
    A[r * lda + c] *= __ldg(s + c);
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void fill_uniform_kernel(int64_t nrows, int ncols, double* __restrict__ A, int64_t lda, uint64_t seed,
                                    int64_t col_offset, int64_t row_offset) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrows * ncols) return;
    const int64_t r = idx / ncols;
    const int64_t c = idx % ncols + col_offset;
    const uint64_t hsh = splitmix64(splitmix64(seed ^ (uint64_t)(r + row_offset) * 0x100000001B3ull) + (uint64_t)c);
    A[r * lda + (c - col_offset)] = (double)(hsh >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

__global__ void gather_rows_kernel(int64_t nrows, int ncols, const double* __restrict__ in, int64_t ldin,
                                   const int* __restrict__ perm, int block, double* __restrict__ out, int64_t ldout) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrows * ncols) return;
    const int64_t r = idx / ncols;
    const int c = (int)(idx % ncols);
    const int64_t src = (int64_t)__ldg(perm + r / block) * block + (r % block);
    out[r * ldout + c] = __ldg(in + src * ldin + c);
}

}  // namespace rvgp

using namespace rvgp;

extern "C" int64_t rvgp_dgemm_workspace_bytes(int m, int n, int split_k) {
    return split_k > 1 ? (int64_t)split_k * m * n * (int64_t)sizeof(double) : 0;
}

namespace rvgp {
int dgemm_launch(Handle* h, int m, int n, int64_t k, double alpha, const double* A, int64_t lda, int a_kmajor,
                 const double* B, int64_t ldb, int b_kmajor, const double* scale_k, double beta, double* C, int64_t ldc,
                 int split_k, double* workspace, int lower_only) {
    RVGP_REQUIRE(h, m >= 0 && n >= 0 && k >= 0 && split_k >= 1, "dgemm: bad sizes");
    RVGP_REQUIRE(h, split_k == 1 || workspace != nullptr, "dgemm: split_k > 1 needs a workspace");
    if (m == 0 || n == 0) return RVGP_OK;
    // bit 1 of lower_only: C aliases A (in-place right-multiplication of a row panel).  Safe only when ONE CTA column covers all
    // of n, so that a CTA has consumed every element of its own rows before it stores them: the 128 x 64-tile kernel, n <= 64.
    const bool in_place = (lower_only & 2) != 0;
    lower_only &= 1;
    RVGP_REQUIRE(h, !in_place || (n <= DBN && split_k == 1), "dgemm: in-place needs n <= 64 and no split-K");
    // pipelined kernel: needs 16-byte aligned 2-element chunks along the contiguous index of every operand (and of C / ws)
    const bool pipe_ok = !in_place && h->dgemm_dmma >= 2 && scale_k == nullptr && !b_kmajor && (n % 2 == 0) && (lda % 2 == 0) &&
                         (ldb % 2 == 0) && (ldc % 2 == 0) && (a_kmajor ? (k % 2 == 0) : (m % 2 == 0)) &&
                         ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C) |
                           reinterpret_cast<uintptr_t>(workspace)) % 16 == 0);
    const bool small_ok = !in_place && h->dgemm_dmma >= 2 && scale_k == nullptr && split_k == 1 &&
                          (int64_t)cdiv(m, BM) * cdiv(n, DBN) * 4 < h->sm_count;
    if (small_ok) {
        dim3 grid(cdiv(n, SBN), cdiv(m, SBM), 1);
#define RVGP_GEMM(AK, BKM) dgemm_small_kernel<AK, BKM><<<grid, 128, 0, h->stream>>>(m, n, k, alpha, A, lda, B, ldb, C, ldc, beta, lower_only)
        if (a_kmajor && b_kmajor) RVGP_GEMM(true, true);
        else if (a_kmajor && !b_kmajor) RVGP_GEMM(true, false);
        else if (!a_kmajor && b_kmajor) RVGP_GEMM(false, true);
        else RVGP_GEMM(false, false);
#undef RVGP_GEMM
    } else if (pipe_ok) {
        if (!h->dgemm_pipe_attr) {
            RVGP_CUDA_OK(h, cudaFuncSetAttribute(dgemm_dmma_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_SMEM));
            RVGP_CUDA_OK(h, cudaFuncSetAttribute(dgemm_dmma_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_SMEM));
            h->dgemm_pipe_attr = 1;
        }
        dim3 grid(cdiv(n, DBN), cdiv(m, BM), split_k);
        if (a_kmajor) dgemm_dmma_pipe_kernel<true><<<grid, 256, PIPE_SMEM, h->stream>>>(m, n, k, alpha, A, lda, B, ldb, C, ldc, split_k, workspace, beta, lower_only);
        else dgemm_dmma_pipe_kernel<false><<<grid, 256, PIPE_SMEM, h->stream>>>(m, n, k, alpha, A, lda, B, ldb, C, ldc, split_k, workspace, beta, lower_only);
    } else if (h->dgemm_dmma) {
        dim3 grid(cdiv(n, DBN), cdiv(m, BM), split_k);
#define RVGP_GEMM(AK, BKM) do { if (scale_k) dgemm_dmma_kernel<AK, BKM, true><<<grid, 256, 0, h->stream>>>(m, n, k, alpha, A, lda, B, ldb, scale_k, C, ldc, split_k, workspace, beta, lower_only); else dgemm_dmma_kernel<AK, BKM, false><<<grid, 256, 0, h->stream>>>(m, n, k, alpha, A, lda, B, ldb, scale_k, C, ldc, split_k, workspace, beta, lower_only); } while (0)
        if (a_kmajor && b_kmajor) RVGP_GEMM(true, true);
        else if (a_kmajor && !b_kmajor) RVGP_GEMM(true, false);
        else if (!a_kmajor && b_kmajor) RVGP_GEMM(false, true);
        else RVGP_GEMM(false, false);
#undef RVGP_GEMM
    } else {
        dim3 grid(cdiv(n, BN), cdiv(m, BM), split_k);
#define RVGP_GEMM(AK, BKM) do { if (scale_k) dgemm_kernel<AK, BKM, true><<<grid, 256, 0, h->stream>>>(m, n, k, alpha, A, lda, B, ldb, scale_k, C, ldc, split_k, workspace, beta, lower_only); else dgemm_kernel<AK, BKM, false><<<grid, 256, 0, h->stream>>>(m, n, k, alpha, A, lda, B, ldb, scale_k, C, ldc, split_k, workspace, beta, lower_only); } while (0)
        if (a_kmajor && b_kmajor) RVGP_GEMM(true, true);
        else if (a_kmajor && !b_kmajor) RVGP_GEMM(true, false);
        else if (!a_kmajor && b_kmajor) RVGP_GEMM(false, true);
        else RVGP_GEMM(false, false);
#undef RVGP_GEMM
    }
    RVGP_LAUNCH_OK(h, "dgemm_kernel");
    if (split_k > 1) {
        const int64_t tot = (int64_t)m * n;
        splitk_reduce_kernel<<<cdiv(tot, 256), 256, 0, h->stream>>>(m, n, split_k, alpha, workspace, C, ldc, beta);
        RVGP_LAUNCH_OK(h, "splitk_reduce_kernel");
    }
    return RVGP_OK;
}
}  // namespace rvgp

extern "C" int rvgp_dgemm_f64(rvgp_handle_t hh, int m, int n, int64_t k, double alpha, const double* A, int64_t lda,
                              int a_kmajor, const double* B, int64_t ldb, int b_kmajor, const double* scale_k,
                              double* C, int64_t ldc, int split_k, double* workspace) {
    return dgemm_launch(H(hh), m, n, k, alpha, A, lda, a_kmajor, B, ldb, b_kmajor, scale_k, 0.0, C, ldc, split_k,
                        workspace, 0);
}

// Same as rvgp_dgemm_f64 but only the tiles that intersect the LOWER triangle of C are computed (C symmetric by
// construction: Gram matrices V^T V, V^T A V); the strict upper triangle of C is left untouched / undefined.
extern "C" int rvgp_dgemm_lower_f64(rvgp_handle_t hh, int m, int n, int64_t k, double alpha, const double* A, int64_t lda,
                                    int a_kmajor, const double* B, int64_t ldb, int b_kmajor, double* C, int64_t ldc,
                                    int split_k, double* workspace) {
    return dgemm_launch(H(hh), m, n, k, alpha, A, lda, a_kmajor, B, ldb, b_kmajor, nullptr, 0.0, C, ldc, split_k, workspace, 1);
}

// C = alpha * op(A) op(B) + beta * C   (same layout flags as rvgp_dgemm_f64; no split-K)
extern "C" int rvgp_dgemm_acc_f64(rvgp_handle_t hh, int m, int n, int64_t k, double alpha, const double* A, int64_t lda,
                                  int a_kmajor, const double* B, int64_t ldb, int b_kmajor, double beta, double* C,
                                  int64_t ldc) {
    return dgemm_launch(H(hh), m, n, k, alpha, A, lda, a_kmajor, B, ldb, b_kmajor, nullptr, beta, C, ldc, 1, nullptr, 0);
}

extern "C" int64_t rvgp_coldot_workspace_bytes(int64_t nrows, int ncols) {
    return (int64_t)cdiv(nrows, RED_ROWS_PER_CTA) * ncols * (int64_t)sizeof(double);
}

static int colreduce(Handle* h, int mode, int64_t nrows, int ncols, const double* A, int64_t lda, const double* B,
                     int64_t ldb, const double* theta, double* out, double* ws) {
    RVGP_REQUIRE(h, ncols >= 1 && nrows >= 0 && ws != nullptr, "colreduce: bad args");
    const int nslabs = cdiv(nrows, RED_ROWS_PER_CTA);
    if (nslabs > 0) {
        const dim3 grid(nslabs, cdiv(ncols, 32));
        if (mode == 0) colreduce_stage1<0><<<grid, 256, 0, h->stream>>>(nrows, ncols, A, lda, B, ldb, theta, ws);
        else colreduce_stage1<1><<<grid, 256, 0, h->stream>>>(nrows, ncols, A, lda, B, ldb, theta, ws);
        RVGP_LAUNCH_OK(h, "colreduce_stage1");
    }
    colreduce_stage2<<<cdiv(ncols, 128), 128, 0, h->stream>>>(nslabs, ncols, ws, out);
    RVGP_LAUNCH_OK(h, "colreduce_stage2");
    return RVGP_OK;
}

extern "C" int rvgp_coldot_f64(rvgp_handle_t hh, int64_t nrows, int ncols, const double* A, int64_t lda,
                               const double* B, int64_t ldb, double* out, double* workspace) {
    return colreduce(H(hh), 0, nrows, ncols, A, lda, B, ldb, nullptr, out, workspace);
}

extern "C" int rvgp_resid_sq_f64(rvgp_handle_t hh, int64_t nrows, int ncols, const double* W, int64_t ldw,
                                 const double* V, int64_t ldv, const double* theta, double* out, double* workspace) {
    return colreduce(H(hh), 1, nrows, ncols, W, ldw, V, ldv, theta, out, workspace);
}

extern "C" int rvgp_colscale_f64(rvgp_handle_t hh, int64_t nrows, int ncols, double* A, int64_t lda, const double* s) {
    Handle* h = H(hh);
    const int64_t tot = nrows * ncols;
    if (tot == 0) return RVGP_OK;
    colscale_kernel<<<cdiv(tot, 256), 256, 0, h->stream>>>(nrows, ncols, A, lda, s);
    RVGP_LAUNCH_OK(h, "colscale_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_fill_uniform_f64(rvgp_handle_t hh, int64_t nrows, int ncols, double* A, int64_t lda, uint64_t seed,
                                     int64_t col_offset, int64_t row_offset) {
    Handle* h = H(hh);
    const int64_t tot = nrows * ncols;
    if (tot == 0) return RVGP_OK;
    fill_uniform_kernel<<<cdiv(tot, 256), 256, 0, h->stream>>>(nrows, ncols, A, lda, seed, col_offset, row_offset);
    RVGP_LAUNCH_OK(h, "fill_uniform_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_gather_rows_f64(rvgp_handle_t hh, int64_t nrows, int ncols, const double* in, int64_t ldin,
                                    const int32_t* perm, int block, double* out, int64_t ldout) {
    Handle* h = H(hh);
    const int64_t tot = nrows * ncols;
    if (tot == 0) return RVGP_OK;
    gather_rows_kernel<<<cdiv(tot, 256), 256, 0, h->stream>>>(nrows, ncols, in, ldin, perm, block, out, ldout);
    RVGP_LAUNCH_OK(h, "gather_rows_kernel");
    return RVGP_OK;
}

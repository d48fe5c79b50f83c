// K14: blocked right-looking Cholesky + triangular solves (FP64), K13 helper K_diag, small GP utilities.
//
// Replaces what GPflow/TensorFlow do for the reference's GPR (RVGP/main.py:55-58,77,80,111): tf.linalg.cholesky
// of K + sigma^2 I, tf.linalg.triangular_solve, sum(log(diag L)); and ManifoldKernel.K_diag (RVGP/kernels.py:63-67)
// WITHOUT forming the full N* x N* matrix the reference builds just to read its diagonal.
//
// potrf: for each 64-wide panel: (1) one CTA factors the diagonal block in shared memory and also emits its
// inverse, (2) the panel below is multiplied by that inverse (dgemm), (3) the trailing lower triangle gets the
// SYRK update as a tile-skipping dgemm.  FP64-FMA-pipe bound: M^3/3 flop.
#include "common.cuh"

namespace rvgp {

int dgemm_launch(Handle* h, int m, int n, int64_t k, double alpha, const double* A, int64_t lda, int a_kmajor,
                 const double* B, int64_t ldb, int b_kmajor, const double* scale_k, double beta, double* C, int64_t ldc,
                 int split_k, double* workspace, int lower_only);

constexpr int NB = 64;

// Factor the nb x nb diagonal block at A (lower, in place; the strict upper part is left untouched) and write
// inv(L_block) (lower, row-major NB x NB, zero above the diagonal, identity in the padding) to Linv.  flag |= 1 on a
// non-positive pivot.
// Round 2: register-resident, ONE __syncthreads per column (the first version kept the block in shared memory, used three
// barriers per column plus a serial 64-step substitution per column of the inverse: 93 us per block, 800 blocks per C4 fit;
// this one: 34 us).
// Thread (i = t >> 2, cq = t & 3) owns row i's entries of columns cq, cq + 4, ... of the block (a[]) and of R (r[]), the running
// right-hand side of L X = I.  At column j its owners publish column j of the block and row j of R; after the barrier every
// thread applies  A[i][c] -= L[i][j] L[c][j]  (j < c <= i)  and  R[i][c] -= L[i][j] X[j][c]  (c <= j < i)  to its registers.
// The buffers alternate, so a thread that runs ahead writes the other copy and one barrier per column is enough.
__global__ void __launch_bounds__(256)
potf2_inv_kernel(double* __restrict__ A, int64_t lda, int nb, double* __restrict__ Linv, int* __restrict__ flag) {
    __shared__ double colbuf[2][NB];
    __shared__ double rowbuf[2][NB];
    const int t = threadIdx.x, i = t >> 2, cq = t & 3;
    double a[16], r[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int c = cq + 4 * q;
        a[q] = (i < nb && c < nb) ? ((c <= i) ? A[(int64_t)i * lda + c] : 0.0) : ((i == c) ? 1.0 : 0.0);
        r[q] = (i == c) ? 1.0 : 0.0;
    }
    bool bad = false;
    // Fully unrolled: every register index is a compile-time constant and the q-ranges are pruned statically.  Measured on
    // B200: 34 us per block (the ~10k-instruction body runs at instruction-fetch speed on its single CTA).  Two variants
    // were tried and dropped: a run-time loop over groups of four columns with select chains for the register indices
    // (90 us: ~700 issued instructions per column), and rsqrt instead of sqrt + reciprocal (2e-10 relative error in the
    // log-likelihood of ill-conditioned dense kernels, over the 1e-10 parity tolerance).
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const int jq = j >> 2, jc = j & 3, pb = j & 1;
        if (cq == jc) colbuf[pb][i] = a[jq];
        if (i == j) {
#pragma unroll
            for (int q = 0; q <= jq; ++q) rowbuf[pb][cq + 4 * q] = r[q];
        }
        __syncthreads();
        double p = colbuf[pb][j];
        if (!(p > 0.0)) { bad = true; p = 1.0; }
        const double d = sqrt(p), inv = 1.0 / d;
        const double ci = colbuf[pb][i] * inv;                      // L[i][j]
        if (i > j) {
#pragma unroll
            for (int q = jq; q < 16; ++q) {
                const int c = cq + 4 * q;
                if (c > j && c <= i) a[q] = fma(-ci, colbuf[pb][c] * inv, a[q]);
            }
#pragma unroll
            for (int q = 0; q <= jq; ++q) {
                const int c = cq + 4 * q;
                if (c <= j) r[q] = fma(-ci, rowbuf[pb][c] * inv, r[q]);
            }
            if (cq == jc) a[jq] = ci;
        } else if (i == j) {
            if (cq == jc) a[jq] = d;
#pragma unroll
            for (int q = 0; q <= jq; ++q) r[q] *= inv;              // X[j][c] = R[j][c] / L[j][j]
        }
    }
    if (bad && t == 0) atomicOr(flag, 1);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int c = cq + 4 * q;
        if (i < nb && c <= i) A[(int64_t)i * lda + c] = a[q];
        Linv[i * NB + c] = (c <= i) ? r[q] : 0.0;
    }
}

__global__ void add_diag_kernel(double* __restrict__ A, int64_t lda, int n, double v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[(int64_t)i * lda + i] += v;
}

// out[0] = sum_i log(A[i,i])   (single CTA, fixed order)
__global__ void __launch_bounds__(256) logdiag_sum_kernel(const double* __restrict__ A, int64_t lda, int n, double* __restrict__ out) {
    __shared__ double sm[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += log(A[(int64_t)i * lda + i]);
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0];
}

// kernels.py:63-67  K_diag[i] = sum_j S[j] X[i,j]^2
__global__ void kdiag_kernel(const double* __restrict__ X, int64_t ldx, int64_t n, int k, const double* __restrict__ S,
                             double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    double s = 0.0;
    for (int j = lane; j < k; j += 32) { const double v = __ldg(X + row * ldx + j); s = fma(__ldg(S + j) * v, v, s); }
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}

}  // namespace rvgp

using namespace rvgp;

extern "C" int64_t rvgp_potrf_workspace_bytes(int n) {
    const int64_t nblk = (n + NB - 1) / NB;
    return (int64_t)n * NB * 8 + nblk * NB * NB * 8;   // panel temp + inverses of the diagonal blocks
}

// In-place lower Cholesky of the symmetric positive definite A (n x n, row-major, only the lower triangle is
// read or written).  workspace: rvgp_potrf_workspace_bytes(n); on return its tail holds inv(L_jj) of every
// diagonal block (used by rvgp_trsm_f64).  flag (device int32, zeroed here): bit0 = not positive definite.
extern "C" int rvgp_potrf_f64(rvgp_handle_t hh, double* A, int64_t lda, int n, int32_t* flag, void* workspace,
                              int64_t workspace_bytes) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, n >= 0 && lda >= n, "potrf: bad sizes");
    if (rvgp_potrf_workspace_bytes(n) > workspace_bytes) return set_error(h, RVGP_ERR_CAPACITY, "potrf: workspace too small%s%s");
    double* T = (double*)workspace;
    double* dinv = T + (int64_t)n * NB;
    RVGP_CUDA_OK(h, cudaMemsetAsync(flag, 0, sizeof(int), h->stream));
    // Two-level blocking: inside an outer panel of NBO = 256 columns the 64-wide steps only update the rest of THAT panel
    // (left-looking within the panel); the trailing matrix then gets ONE rank-256 SYRK per outer panel, which keeps the
    // big update at K = 256 (compute-bound) instead of K = 64 (12 -> ~20 TFLOP/s at M = 32k).
    constexpr int NBO = 256;
    for (int J0 = 0; J0 < n; J0 += NBO) {
        const int W = (n - J0 < NBO) ? n - J0 : NBO;
        for (int j0 = J0; j0 < J0 + W; j0 += NB) {
            const int b = j0 / NB;
            const int nb = (J0 + W - j0 < NB) ? J0 + W - j0 : NB;
            double* Ajj = A + (int64_t)j0 * lda + j0;
            double* Linv = dinv + (int64_t)b * NB * NB;
            potf2_inv_kernel<<<1, 256, 0, h->stream>>>(Ajj, lda, nb, Linv, flag);
            RVGP_LAUNCH_OK(h, "potf2_inv_kernel");
            const int m2 = n - j0 - nb;
            if (m2 > 0) {
                double* A21 = A + (int64_t)(j0 + nb) * lda + j0;
                // A21 <- A21 * inv(L11)^T, IN PLACE (flag 2: one CTA column, every CTA reads only its own 128 rows of A21 and has
                // consumed all of them, K = nb, before its epilogue stores them)
                int rc = dgemm_launch(h, m2, nb, nb, 1.0, A21, lda, 1, Linv, NB, 1, nullptr, 0.0, A21, lda, 1, nullptr, 2);
                if (rc) return rc;
                // update the remaining columns of this outer panel only
                const int wrem = J0 + W - (j0 + nb);
                if (wrem > 0) {
                    double* C = A + (int64_t)(j0 + nb) * lda + (j0 + nb);
                    rc = dgemm_launch(h, m2, wrem, nb, -1.0, A21, lda, 1, A21, lda, 1, nullptr, 1.0, C, lda, 1, nullptr, 0);
                    if (rc) return rc;
                }
            }
        }
        const int mt = n - J0 - W;
        if (mt > 0) {   // trailing rank-W SYRK (lower tiles): A22 -= P P^T, P = A[J0+W:, J0:J0+W]
            double* P = A + (int64_t)(J0 + W) * lda + J0;
            double* A22 = A + (int64_t)(J0 + W) * lda + (J0 + W);
            int rc = dgemm_launch(h, mt, mt, W, -1.0, P, lda, 1, P, lda, 1, nullptr, 1.0, A22, lda, 1, nullptr, 1);
            if (rc) return rc;
        }
    }
    return RVGP_OK;
}

// Solve L X = B (trans = 0) or L^T X = B (trans = 1) in place; L is the rvgp_potrf_f64 output (lower), B is
// (n x nrhs) row-major.  `workspace` must be the SAME buffer rvgp_potrf_f64 used (diagonal-block inverses) and
// additionally needs NB*nrhs doubles of scratch at `scratch`.
extern "C" int rvgp_trsm_f64(rvgp_handle_t hh, const double* L, int64_t ldl, int n, double* B, int64_t ldb, int nrhs, int trans,
                             const void* potrf_workspace, double* scratch) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, n >= 0 && nrhs >= 0 && (trans == 0 || trans == 1), "trsm: bad args");
    if (n == 0 || nrhs == 0) return RVGP_OK;
    const double* dinv = (const double*)potrf_workspace + (int64_t)n * NB;
    const int nblk = (n + NB - 1) / NB;
    for (int s = 0; s < nblk; ++s) {
        const int b = trans ? nblk - 1 - s : s;
        const int j0 = b * NB;
        const int nb = (n - j0 < NB) ? n - j0 : NB;
        const double* Linv = dinv + (int64_t)b * NB * NB;
        double* Bj = B + (int64_t)j0 * ldb;
        // X_j = inv(L_jj) B_j   or   inv(L_jj)^T B_j
        int rc = dgemm_launch(h, nb, nrhs, nb, 1.0, Linv, NB, trans ? 0 : 1, Bj, ldb, 0, nullptr, 0.0, scratch, nrhs, 1, nullptr, 0);
        if (rc) return rc;
        RVGP_CUDA_OK(h, cudaMemcpy2DAsync(Bj, ldb * sizeof(double), scratch, (size_t)nrhs * sizeof(double),
                                          (size_t)nrhs * sizeof(double), (size_t)nb, cudaMemcpyDeviceToDevice, h->stream));
        if (!trans) {
            const int m2 = n - j0 - nb;        // B[j+1:] -= L[j+1:, j] X_j
            if (m2 > 0) {
                rc = dgemm_launch(h, m2, nrhs, nb, -1.0, L + (int64_t)(j0 + nb) * ldl + j0, ldl, 1, Bj, ldb, 0, nullptr, 1.0,
                                  B + (int64_t)(j0 + nb) * ldb, ldb, 1, nullptr, 0);
                if (rc) return rc;
            }
        } else if (j0 > 0) {                   // B[:j] -= L[j, :j]^T X_j
            rc = dgemm_launch(h, j0, nrhs, nb, -1.0, L + (int64_t)j0 * ldl, ldl, 0, Bj, ldb, 0, nullptr, 1.0, B, ldb, 1,
                              nullptr, 0);
            if (rc) return rc;
        }
    }
    return RVGP_OK;
}

extern "C" int rvgp_add_diag_f64(rvgp_handle_t hh, double* A, int64_t lda, int n, double v) {
    Handle* h = H(hh);
    if (n == 0) return RVGP_OK;
    add_diag_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(A, lda, n, v);
    RVGP_LAUNCH_OK(h, "add_diag_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_logdiag_sum_f64(rvgp_handle_t hh, const double* A, int64_t lda, int n, double* out) {
    Handle* h = H(hh);
    logdiag_sum_kernel<<<1, 256, 0, h->stream>>>(A, lda, n, out);
    RVGP_LAUNCH_OK(h, "logdiag_sum_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_kdiag_f64(rvgp_handle_t hh, const double* X, int64_t ldx, int64_t n, int k, const double* S, double* out) {
    Handle* h = H(hh);
    if (n == 0) return RVGP_OK;
    kdiag_kernel<<<cdiv(n * 32, 256), 256, 0, h->stream>>>(X, ldx, n, k, S, out);
    RVGP_LAUNCH_OK(h, "kdiag_kernel");
    return RVGP_OK;
}

// ---- K15b: one rank-k GP evaluation (k > 64) without host round trips ------------------------------------------------------
// The 4-scalar L-BFGS-B loop of train_gp (main.py:87-95) evaluates the rank-k log marginal likelihood ~100 times per fit; each
// evaluation is O(k^3) on resident k x k data, so its cost is launch / synchronisation overhead, not arithmetic.  This entry
// point enqueues the WHOLE evaluation (build B = I + S^1/2 G S^1/2 / noise, Cholesky, three triangular solves, column norms,
// log-determinant) and packs what the host needs into `out`; the caller then makes ONE device->host copy of 2 + 2k doubles.
namespace rvgp {
__global__ void gp_lowrank_build_kernel(int k, const double* __restrict__ G, const double* __restrict__ b,
                                        const double* __restrict__ par, double* __restrict__ B, double* __restrict__ Q,
                                        double* __restrict__ rhs) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)k * k) return;
    const int i = (int)(idx / k), j = (int)(idx - (int64_t)i * k);
    const double noise = par[k];
    const double ri = sqrt(par[i]), rj = sqrt(par[j]);
    const double g = G[idx];
    B[idx] = ((i == j) ? 1.0 : 0.0) + ri * g * rj / noise;
    if (Q) Q[idx] = ri * g;
    if (rhs && j == 0) rhs[i] = ri * b[i];
}

__global__ void gp_lowrank_pack_kernel(int k, const int* __restrict__ flag, const double* __restrict__ logdet_half,
                                       const double* __restrict__ z, const double* __restrict__ qs, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { out[0] = (double)(*flag & 1); out[1] = *logdet_half; }
    if (i < k) { out[2 + i] = z[i]; out[2 + k + i] = qs[i]; }
}

// Forward substitution  Y = L^-1 [S^1/2 G | S^1/2 b]  for FW right-hand-side columns per CTA, all panels in ONE launch (the
// panel-by-panel rvgp_trsm_f64 needs 3 launches per 64 rows and per solve: 72 of the ~110 launches of an evaluation).  The CTA
// keeps its k x FW slice in shared memory, walks the 64-row panels (X_j = inv(L_jj) Y_j with the diagonal-block inverses potrf
// left behind, then Y[below] -= L[below, j] X_j) and finishes with what the evaluation needs: the squared column norms
// qs_c = || L^-1 (S^1/2 G) e_c ||^2 for c < k, and the solved vector itself for the right-hand side S^1/2 b (column k).
constexpr int FW = 8;
__global__ void __launch_bounds__(256)
gp_lowrank_fwd_kernel(int k, const double* __restrict__ L, const double* __restrict__ dinv, const double* __restrict__ G,
                      const double* __restrict__ b, const double* __restrict__ par, double* __restrict__ qs,
                      double* __restrict__ w) {
    extern __shared__ __align__(16) double fwd_smem[];
    double (*Y)[FW] = reinterpret_cast<double (*)[FW]>(fwd_smem);                 // [kpad][FW]
    const int kpad = (k + NB - 1) / NB * NB;
    double* Ds = fwd_smem + (size_t)kpad * FW;                                    // [NB][NB + 1]: inverse of the current diagonal block
    __shared__ __align__(16) double Xs[NB][FW];
    __shared__ double red[32][FW];
    const int t = threadIdx.x;
    const int c0 = blockIdx.x * FW;
    for (int e = t; e < kpad * FW; e += 256) {
        const int i = e / FW, c = e % FW, col = c0 + c;
        double v = 0.0;
        if (i < k && col <= k) v = sqrt(par[i]) * ((col < k) ? G[(int64_t)i * k + col] : b[i]);
        Y[i][c] = v;
    }
    for (int j0 = 0; j0 < k; j0 += NB) {
        const int nb = (k - j0 < NB) ? k - j0 : NB;
        const double* Dj = dinv + (int64_t)(j0 / NB) * NB * NB;
#pragma unroll
        for (int q = 0; q < NB * NB / 256; ++q) {                                 // 16 independent loads per thread in flight
            const int e = t + 256 * q;
            Ds[(e >> 6) * (NB + 1) + (e & 63)] = __ldg(Dj + e);
        }
        __syncthreads();                                                          // Ds, and the previous panel's updates of Y
        {   // X_j = inv(L_jj) Y_j : thread (r, column pair)
            const int r = t >> 2, cp = (t & 3) * 2;
            double x0 = 0.0, x1 = 0.0;
            if (r < nb) {
                for (int m = 0; m <= r; ++m) {
                    const double l = Ds[r * (NB + 1) + m];
                    x0 = fma(l, Y[j0 + m][cp], x0);
                    x1 = fma(l, Y[j0 + m][cp + 1], x1);
                }
            }
            Xs[r][cp] = x0; Xs[r][cp + 1] = x1;
            __syncthreads();                                                      // every read of Y_j and Ds is done
            if (r < nb) { Y[j0 + r][cp] = x0; Y[j0 + r][cp + 1] = x1; }
        }
        // rows below the panel: one row per thread, the row's 64 entries of L fetched 16 at a time (the loads of a batch are
        // independent and in flight together: with one load per multiply-add this loop ran at L2 latency, 125 us per launch)
        for (int i = j0 + nb + t; i < k; i += 256) {
            double acc[FW];
#pragma unroll
            for (int c = 0; c < FW; ++c) acc[c] = Y[i][c];
            const double* Li = L + (int64_t)i * k + j0;
            for (int mb = 0; mb < nb; mb += 16) {
                double l[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) l[q] = (mb + q < nb) ? __ldg(Li + mb + q) : 0.0;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int m = (mb + q < NB) ? mb + q : NB - 1;
                    const double2 x01 = *reinterpret_cast<const double2*>(&Xs[m][0]);
                    const double2 x23 = *reinterpret_cast<const double2*>(&Xs[m][2]);
                    const double2 x45 = *reinterpret_cast<const double2*>(&Xs[m][4]);
                    const double2 x67 = *reinterpret_cast<const double2*>(&Xs[m][6]);
                    acc[0] = fma(-l[q], x01.x, acc[0]); acc[1] = fma(-l[q], x01.y, acc[1]);
                    acc[2] = fma(-l[q], x23.x, acc[2]); acc[3] = fma(-l[q], x23.y, acc[3]);
                    acc[4] = fma(-l[q], x45.x, acc[4]); acc[5] = fma(-l[q], x45.y, acc[5]);
                    acc[6] = fma(-l[q], x67.x, acc[6]); acc[7] = fma(-l[q], x67.y, acc[7]);
                }
            }
#pragma unroll
            for (int c = 0; c < FW; ++c) Y[i][c] = acc[c];
        }
        // (the barrier at the top of the next panel orders these updates, Xs and Ds)
    }
    __syncthreads();
    {   // squared column norms, fixed order
        const int c = t & (FW - 1), g = t >> 3;                                   // 32 row groups
        double sacc = 0.0;
        for (int i = g; i < k; i += 32) sacc = fma(Y[i][c], Y[i][c], sacc);
        red[g][c] = sacc;
        __syncthreads();
        if (t < FW) {
            double tot = 0.0;
            for (int gg = 0; gg < 32; ++gg) tot += red[gg][t];
            if (c0 + t < k) qs[c0 + t] = tot;
        }
    }
    if (c0 <= k && k < c0 + FW) {
        const int c = k - c0;
        for (int i = t; i < k; i += 256) w[i] = Y[i][c];
    }
}

// Back substitution  z = L^-T w  (single CTA: k^2 / 2 multiply-adds), sum log L_ii, and the packing of the evaluation's
// read-back [not-SPD flag, sum log L_ii, z (k), qs (k)].
__global__ void __launch_bounds__(256)
gp_lowrank_finish_kernel(int k, const double* __restrict__ L, const double* __restrict__ dinv, const double* __restrict__ w,
                         const int* __restrict__ flag, const double* __restrict__ qs, double* __restrict__ out) {
    extern __shared__ double fin_smem[];                                          // [kpad] running right-hand side / solution
    __shared__ double zs[NB];
    __shared__ double red[256];
    __shared__ double Ds[NB * (NB + 1)];
    const int t = threadIdx.x;
    const int kpad = (k + NB - 1) / NB * NB;
    for (int i = t; i < kpad; i += 256) fin_smem[i] = (i < k) ? w[i] : 0.0;
    const int nblk = kpad / NB;
    for (int bj = nblk - 1; bj >= 0; --bj) {
        const int j0 = bj * NB;
        const int nb = (k - j0 < NB) ? k - j0 : NB;
        const double* Dj = dinv + (int64_t)bj * NB * NB;
#pragma unroll
        for (int q = 0; q < NB * NB / 256; ++q) {
            const int e = t + 256 * q;
            Ds[(e >> 6) * (NB + 1) + (e & 63)] = __ldg(Dj + e);
        }
        __syncthreads();
        {   // z_j = inv(L_jj)^T w_j : 4 threads per entry, interleaved over m
            const int r = t >> 2, part = t & 3;
            double sacc = 0.0;
            if (r < nb)
                for (int m = r + part; m < nb; m += 4) sacc = fma(Ds[m * (NB + 1) + r], fin_smem[j0 + m], sacc);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
            if (part == 0) zs[r] = sacc;
        }
        __syncthreads();
        if (t < nb) fin_smem[j0 + t] = zs[t];
        for (int i = t; i < j0; i += 256) {                                       // w[:j0] -= L[j0:j0+nb, :j0]^T z_j
            double a = fin_smem[i];
            const double* Lc = L + (int64_t)j0 * k + i;
            for (int mb = 0; mb < nb; mb += 16) {
                double l[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) l[q] = (mb + q < nb) ? __ldg(Lc + (int64_t)(mb + q) * k) : 0.0;
#pragma unroll
                for (int q = 0; q < 16; ++q) a = fma(-l[q], zs[(mb + q < NB) ? mb + q : NB - 1], a);
            }
            fin_smem[i] = a;
        }
        __syncthreads();
    }
    double ld = 0.0;
    for (int i = t; i < k; i += 256) ld += log(L[(int64_t)i * k + i]);
    red[t] = ld;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (t < o) red[t] += red[t + o];
        __syncthreads();
    }
    if (t == 0) { out[0] = (double)(*flag & 1); out[1] = red[0]; }
    for (int i = t; i < k; i += 256) { out[2 + i] = fin_smem[i]; out[2 + k + i] = qs[i]; }
}
}  // namespace rvgp

extern "C" int64_t rvgp_coldot_workspace_bytes(int64_t nrows, int ncols);
extern "C" int rvgp_coldot_f64(rvgp_handle_t hh, int64_t nrows, int ncols, const double* A, int64_t lda, const double* B,
                               int64_t ldb, double* out, double* workspace);

// G (k x k), b (k): resident Gram data of the fit (Phi^T Phi, Phi^T y).  par: [S (k), noise] on the device.
// work: 2 k^2 + 3 k + 2 doubles + rvgp_gp_lowrank_eval_workspace_bytes(k) in total, laid out by this function.
// out (2 + 2k doubles): [not-SPD flag, sum log L_ii, z = B^-1 (S^1/2 b), qs_j = || L^-1 (S^1/2 G) e_j ||^2].
extern "C" int64_t rvgp_gp_lowrank_eval_workspace_bytes(int k) {
    const int64_t dbl = 2 * (int64_t)k * k + 3 * (int64_t)k + 16 + 64 * (int64_t)k;
    return dbl * (int64_t)sizeof(double) + rvgp_potrf_workspace_bytes(k) + rvgp_coldot_workspace_bytes(k, k) + 256;
}

static int gp_lowrank_eval_enqueue(rvgp_handle_t hh, int k, const double* G, const double* b, const double* par, double* out,
                                   void* workspace) {
    Handle* h = H(hh);
    double* B = (double*)workspace;
    double* Q = B + (int64_t)k * k;
    double* rhs = Q + (int64_t)k * k;
    double* qs = rhs + k;
    double* logdet = qs + k;                       // 1 double (+ padding)
    int32_t* flag = (int32_t*)(logdet + 8);
    double* scratch = logdet + 16;                 // 64 * k doubles (rvgp_trsm_f64)
    char* tail = (char*)(scratch + 64 * (int64_t)k);
    tail = (char*)(((uintptr_t)tail + 255) / 256 * 256);
    void* potrf_ws = tail;
    const int64_t pwb = rvgp_potrf_workspace_bytes(k);
    double* cd_ws = (double*)(tail + (pwb + 255) / 256 * 256);
    const int kpad = (k + NB - 1) / NB * NB;
    const size_t fwd_smem = ((size_t)kpad * FW + NB * (NB + 1)) * sizeof(double);
    if (fwd_smem <= 200 * 1024) {
        // fused solves: build, Cholesky, ONE forward kernel over column slices of [S^1/2 G | S^1/2 b], ONE finishing kernel
        gp_lowrank_build_kernel<<<cdiv((int64_t)k * k, 256), 256, 0, h->stream>>>(k, G, b, par, B, nullptr, nullptr);
        RVGP_LAUNCH_OK(h, "gp_lowrank_build_kernel");
        int rc;
        if ((rc = rvgp_potrf_f64(hh, B, k, k, flag, potrf_ws, pwb))) return rc;
        const double* dinv = (const double*)potrf_ws + (int64_t)k * NB;
        if (fwd_smem > 32 * 1024)                              // + 6 KB of static shared memory: past the 48 KB default
            RVGP_CUDA_OK(h, cudaFuncSetAttribute(gp_lowrank_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem));
        gp_lowrank_fwd_kernel<<<cdiv(k + 1, FW), 256, fwd_smem, h->stream>>>(k, B, dinv, G, b, par, qs, rhs);
        RVGP_LAUNCH_OK(h, "gp_lowrank_fwd_kernel");
        gp_lowrank_finish_kernel<<<1, 256, (size_t)kpad * sizeof(double), h->stream>>>(k, B, dinv, rhs, flag, qs, out);
        RVGP_LAUNCH_OK(h, "gp_lowrank_finish_kernel");
        return RVGP_OK;
    }
    gp_lowrank_build_kernel<<<cdiv((int64_t)k * k, 256), 256, 0, h->stream>>>(k, G, b, par, B, Q, rhs);
    RVGP_LAUNCH_OK(h, "gp_lowrank_build_kernel");
    int rc;
    if ((rc = rvgp_potrf_f64(hh, B, k, k, flag, potrf_ws, pwb))) return rc;
    if ((rc = rvgp_trsm_f64(hh, B, k, k, rhs, 1, 1, 0, potrf_ws, scratch))) return rc;
    if ((rc = rvgp_trsm_f64(hh, B, k, k, rhs, 1, 1, 1, potrf_ws, scratch))) return rc;
    if ((rc = rvgp_trsm_f64(hh, B, k, k, Q, k, k, 0, potrf_ws, scratch))) return rc;
    if ((rc = rvgp_coldot_f64(hh, k, k, Q, k, Q, k, qs, cd_ws))) return rc;
    if ((rc = rvgp_logdiag_sum_f64(hh, B, k, k, logdet))) return rc;
    gp_lowrank_pack_kernel<<<cdiv(k, 256), 256, 0, h->stream>>>(k, flag, logdet, rhs, qs, out);
    RVGP_LAUNCH_OK(h, "gp_lowrank_pack_kernel");
    return RVGP_OK;
}

// The evaluation is ~30 small launches (64-wide Cholesky panels, then the two fused solve kernels; ~110 before the solves were
// fused, and still that many for k > 3200 where the forward kernel's slice no longer fits in shared memory): it is captured into
// a CUDA graph on first use and replayed with ONE cudaGraphLaunch while (k, G, b, par, out, workspace) stay the same -- which
// they do for the ~100 evaluations of a fit.  The legacy default stream cannot be captured, so a handle that follows it uses
// a BLOCKING side stream (implicitly ordered with the legacy stream on both sides).  Any capture failure falls back to eager
// launches for the rest of the handle's life.
extern "C" int rvgp_gp_lowrank_eval_f64(rvgp_handle_t hh, int k, const double* G, const double* b, const double* par,
                                        double* out, void* workspace, int64_t workspace_bytes) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, k >= 1, "gp_lowrank_eval: k >= 1");
    if (rvgp_gp_lowrank_eval_workspace_bytes(k) > workspace_bytes)
        return set_error(h, RVGP_ERR_CAPACITY, "gp_lowrank_eval: workspace too small%s%s");
    if (h->gp_graph_off) return gp_lowrank_eval_enqueue(hh, k, G, b, par, out, workspace);
    const void* key[6] = {(const void*)(intptr_t)k, G, b, par, out, workspace};
    cudaStream_t user = h->stream;
    cudaStream_t run = user;
    if (user == nullptr || user == cudaStreamLegacy) {
        if (h->gp_stream == nullptr && cudaStreamCreate(&h->gp_stream) != cudaSuccess) { h->gp_graph_off = 1; cudaGetLastError(); }
        if (h->gp_graph_off) return gp_lowrank_eval_enqueue(hh, k, G, b, par, out, workspace);
        run = h->gp_stream;
    }
    if (h->gp_graph != nullptr && memcmp(key, h->gp_graph_key, sizeof(key)) != 0) {
        cudaGraphExecDestroy(h->gp_graph);
        h->gp_graph = nullptr;
    }
    if (h->gp_graph == nullptr) {
        const unsigned long long l0 = h->launches;
        cudaGraph_t graph = nullptr;
        // warm the lazily loaded kernels / attributes OUTSIDE the capture (first evaluation of a fit runs eagerly)
        h->stream = run;
        int rc = gp_lowrank_eval_enqueue(hh, k, G, b, par, out, workspace);
        if (rc) { h->stream = user; return rc; }
        bool ok = cudaStreamBeginCapture(run, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            rc = gp_lowrank_eval_enqueue(hh, k, G, b, par, out, workspace);
            ok = (cudaStreamEndCapture(run, &graph) == cudaSuccess) && rc == RVGP_OK && graph != nullptr;
        }
        h->stream = user;
        if (ok) ok = cudaGraphInstantiate(&h->gp_graph, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        if (!ok) {
            cudaGetLastError();
            h->gp_graph = nullptr;
            h->gp_graph_off = 1;                   // the eager evaluation above already produced this call's result
            return RVGP_OK;
        }
        memcpy(h->gp_graph_key, key, sizeof(key));
        h->gp_graph_launches = (h->launches - l0) / 2;
        return RVGP_OK;                            // result of this call: the eager run above
    }
    RVGP_CUDA_OK(h, cudaGraphLaunch(h->gp_graph, run));
    h->launches += h->gp_graph_launches;
    return RVGP_OK;
}

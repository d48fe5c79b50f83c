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
// inv(L_block) (lower, row-major nb x NB, zero above the diagonal) to Linv.  flag |= 1 on a non-positive pivot.
__global__ void __launch_bounds__(256)
potf2_inv_kernel(double* __restrict__ A, int64_t lda, int nb, double* __restrict__ Linv, int* __restrict__ flag) {
    extern __shared__ double potf2_smem[];
    double (*Ls)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(potf2_smem);
    double (*Xs)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(potf2_smem + NB * (NB + 1));
    const int t = threadIdx.x;
    for (int e = t; e < NB * NB; e += 256) {
        const int i = e / NB, j = e % NB;
        Ls[i][j] = (i < nb && j <= i) ? A[(int64_t)i * lda + j] : ((i == j) ? 1.0 : 0.0);
        Xs[i][j] = 0.0;
    }
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        if (t == 0) {
            double p = Ls[j][j];
            if (!(p > 0.0)) { atomicOr(flag, 1); p = 1.0; }
            Ls[j][j] = sqrt(p);
        }
        __syncthreads();
        const double djj = Ls[j][j];
        for (int i = j + 1 + t; i < nb; i += 256) Ls[i][j] /= djj;
        __syncthreads();
        const int rem = nb - j - 1;
        for (int e = t; e < rem * rem; e += 256) {
            const int i = j + 1 + e / rem, k = j + 1 + e % rem;
            if (k <= i) Ls[i][k] = fma(-Ls[i][j], Ls[k][j], Ls[i][k]);
        }
        __syncthreads();
    }
    // inverse by forward substitution, one column per thread
    if (t < nb) {
        const int c = t;
        for (int i = c; i < nb; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) s = fma(-Ls[i][k], Xs[k][c], s);
            Xs[i][c] = s / Ls[i][i];
        }
    }
    __syncthreads();
    for (int e = t; e < nb * nb; e += 256) {
        const int i = e / nb, j = e % nb;
        if (j <= i) A[(int64_t)i * lda + j] = Ls[i][j];
        Linv[i * NB + j] = Xs[i][j];
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
    const int potf2_smem_bytes = 2 * NB * (NB + 1) * (int)sizeof(double);
    RVGP_CUDA_OK(h, cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, potf2_smem_bytes));
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
            potf2_inv_kernel<<<1, 256, potf2_smem_bytes, h->stream>>>(Ajj, lda, nb, Linv, flag);
            RVGP_LAUNCH_OK(h, "potf2_inv_kernel");
            const int m2 = n - j0 - nb;
            if (m2 > 0) {
                double* A21 = A + (int64_t)(j0 + nb) * lda + j0;
                // T = A21 * inv(L11)^T  (all rows below the diagonal block)
                int rc = dgemm_launch(h, m2, nb, nb, 1.0, A21, lda, 1, Linv, NB, 1, nullptr, 0.0, T, NB, 1, nullptr, 0);
                if (rc) return rc;
                RVGP_CUDA_OK(h, cudaMemcpy2DAsync(A21, lda * sizeof(double), T, NB * sizeof(double), (size_t)nb * sizeof(double),
                                                  (size_t)m2, cudaMemcpyDeviceToDevice, h->stream));
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
    Q[idx] = ri * g;
    if (j == 0) rhs[i] = ri * b[i];
}

__global__ void gp_lowrank_pack_kernel(int k, const int* __restrict__ flag, const double* __restrict__ logdet_half,
                                       const double* __restrict__ z, const double* __restrict__ qs, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { out[0] = (double)(*flag & 1); out[1] = *logdet_half; }
    if (i < k) { out[2 + i] = z[i]; out[2 + k + i] = qs[i]; }
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

// The evaluation is ~110 tiny launches (64-wide Cholesky panels and triangular solves of a k x k matrix): it is captured into
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

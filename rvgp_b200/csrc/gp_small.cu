// K16: fused rank-k GP evaluation for small k (<= 64): log marginal likelihood and its gradient in ONE launch.
//
// The EEG-shaped workload (reference examples/eeg_example/eeg_utils.py:26-35,107-112) runs one GP fit per time frame on a
// fixed eigenbasis with M ~ 150 rows and k ~ 10 eigenpairs; every L-BFGS-B evaluation is tiny, so it is latency-bound.
// Everything the evaluation needs follows from G = Phi^T Phi (k x k), b = Phi^T y, y^T y (see rvgp_b200/gp.py): this kernel
// builds B = I + S^1/2 G S^1/2 / noise, factors it in shared memory, does the triangular solves and assembles
// LML, dLML/dS (k values) and dLML/dnoise -- one CTA per problem, so many frames can be evaluated in one launch.
#include "common.cuh"

namespace rvgp {

constexpr int KS = 64;

// out (per problem): [0] lml, [1] dnoise, [2 .. 2+k) dS, [2+k .. 2+2k) posterior mean weights wbar, then k*k: Q = L_b^-1 S^1/2
__global__ void __launch_bounds__(256)
gp_lowrank_small_kernel(int k, const double* __restrict__ G, int64_t g_stride, const double* __restrict__ b, int64_t b_stride,
                        const double* __restrict__ yy, const double* __restrict__ Mrows, const double* __restrict__ s,
                        int64_t s_stride, const double* __restrict__ noise_p, double* __restrict__ out, int64_t out_stride,
                        int want_predict) {
    extern __shared__ double gp_small_smem[];
    double (*Bm)[KS + 1] = reinterpret_cast<double (*)[KS + 1]>(gp_small_smem);
    double (*Qm)[KS + 1] = reinterpret_cast<double (*)[KS + 1]>(gp_small_smem + KS * (KS + 1));
    double* rs = gp_small_smem + 2 * KS * (KS + 1);
    double *bt = rs + KS, *z = bt + KS, *cvec = z + KS, *Gc = cvec + KS, *wd = Gc + KS;
    __shared__ int bad;
    const int t = threadIdx.x, prob = blockIdx.x;
    G += prob * g_stride; b += prob * b_stride; s += prob * s_stride; out += prob * out_stride;
    const double noise = noise_p[prob], M = Mrows[prob], y2 = yy[prob];
    if (t == 0) bad = 0;
    if (t < k) { rs[t] = sqrt(s[t]); }
    __syncthreads();
    for (int e = t; e < k * k; e += 256) {
        const int i = e / k, j = e % k;
        Bm[i][j] = ((i == j) ? 1.0 : 0.0) + rs[i] * G[i * k + j] * rs[j] / noise;
        Qm[i][j] = rs[i] * G[i * k + j];
    }
    if (t < k) bt[t] = rs[t] * b[t];
    __syncthreads();
    // Cholesky (lower) in place
    for (int j = 0; j < k; ++j) {
        if (t == 0) {
            double p = Bm[j][j];
            if (!(p > 0.0)) { bad = 1; p = 1.0; }
            Bm[j][j] = sqrt(p);
        }
        __syncthreads();
        const double djj = Bm[j][j];
        for (int i = j + 1 + t; i < k; i += 256) Bm[i][j] /= djj;
        __syncthreads();
        const int rem = k - j - 1;
        for (int e = t; e < rem * rem; e += 256) {
            const int i = j + 1 + e / rem, c = j + 1 + e % rem;
            if (c <= i) Bm[i][c] = fma(-Bm[i][j], Bm[c][j], Bm[i][c]);
        }
        __syncthreads();
    }
    // z = B^-1 bt (thread 0: forward + backward substitution, k <= 64) while threads 1..k solve Q = L^-1 (S^1/2 G)
    if (t == 0) {
        for (int i = 0; i < k; ++i) {
            double v = bt[i];
            for (int c = 0; c < i; ++c) v = fma(-Bm[i][c], z[c], v);
            z[i] = v / Bm[i][i];
        }
        for (int i = k - 1; i >= 0; --i) {
            double v = z[i];
            for (int c = i + 1; c < k; ++c) v = fma(-Bm[c][i], z[c], v);
            z[i] = v / Bm[i][i];
        }
    } else if (t >= 32 && t < 32 + k) {
        const int c = t - 32;
        for (int i = 0; i < k; ++i) {
            double v = Qm[i][c];
            for (int r = 0; r < i; ++r) v = fma(-Bm[i][r], Qm[r][c], v);
            Qm[i][c] = v / Bm[i][i];
        }
    }
    __syncthreads();
    if (t < k) {
        cvec[t] = rs[t] * z[t];
        double q2 = 0.0;
        for (int i = 0; i < k; ++i) q2 = fma(Qm[i][t], Qm[i][t], q2);
        wd[t] = (G[t * k + t] - q2 / noise) / noise;
    }
    __syncthreads();
    if (t < k) {
        double v = 0.0;
        for (int j = 0; j < k; ++j) v = fma(G[t * k + j], cvec[j], v);
        Gc[t] = v;
    }
    __syncthreads();
    if (t < k) {
        const double u = (b[t] - Gc[t] / noise) / noise;
        out[2 + t] = 0.5 * u * u - 0.5 * wd[t];
        out[2 + k + t] = cvec[t] / noise;                    // posterior mean weights
    }
    if (t == 0) {
        double btz = 0.0, bc = 0.0, cGc = 0.0, swd = 0.0, ld = 0.0;
        for (int i = 0; i < k; ++i) {
            btz = fma(bt[i], z[i], btz); bc = fma(b[i], cvec[i], bc); cGc = fma(cvec[i], Gc[i], cGc);
            swd = fma(s[i], wd[i], swd); ld += log(Bm[i][i]);
        }
        const double quad = (y2 - btz / noise) / noise;
        const double lml = -0.5 * quad - 0.5 * M * 1.8378770664093453 - 0.5 * M * log(noise) - ld;
        const double aa = (y2 - 2.0 * bc / noise + cGc / (noise * noise)) / (noise * noise);
        const double tr_inv = (M - swd) / noise;
        out[0] = bad ? __longlong_as_double(0x7ff8000000000000ll) : lml;
        out[1] = 0.5 * aa - 0.5 * tr_inv;
    }
    if (want_predict) {
        // Q2 = L^-1 diag(rs): needed for the predictive variance ||Q2 x*||^2
        __syncthreads();
        for (int e = t; e < k * k; e += 256) Qm[e / k][e % k] = (e / k == e % k) ? rs[e / k] : 0.0;
        __syncthreads();
        if (t < k) {
            const int c = t;
            for (int i = 0; i < k; ++i) {
                double v = Qm[i][c];
                for (int r = 0; r < i; ++r) v = fma(-Bm[i][r], Qm[r][c], v);
                Qm[i][c] = v / Bm[i][i];
            }
        }
        __syncthreads();
        for (int e = t; e < k * k; e += 256) out[2 + 2 * k + e] = Qm[e / k][e % k];
    }
}

}  // namespace rvgp

using namespace rvgp;

// nprob problems, one CTA each.  Strides are in doubles between consecutive problems (0 = shared by all problems).
// out per problem: [lml, dLML/dnoise, dLML/dS (k), wbar (k), (want_predict ? Q (k*k row-major) : nothing)].
extern "C" int rvgp_gp_lowrank_small_f64(rvgp_handle_t hh, int nprob, int k, const double* G, int64_t g_stride, const double* b,
                                         int64_t b_stride, const double* yy, const double* Mrows, const double* s,
                                         int64_t s_stride, const double* noise, double* out, int64_t out_stride,
                                         int want_predict) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, k >= 1 && k <= KS, "gp_lowrank_small: k must be in [1,64]");
    RVGP_REQUIRE(h, out_stride >= 2 + 2 * k + (want_predict ? k * k : 0), "gp_lowrank_small: out_stride too small");
    if (nprob <= 0) return RVGP_OK;
    const int smem = (2 * KS * (KS + 1) + 6 * KS) * (int)sizeof(double);
    RVGP_CUDA_OK(h, cudaFuncSetAttribute(gp_lowrank_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    gp_lowrank_small_kernel<<<nprob, 256, smem, h->stream>>>(k, G, g_stride, b, b_stride, yy, Mrows, s, s_stride, noise, out,
                                                           out_stride, want_predict);
    RVGP_LAUNCH_OK(h, "gp_lowrank_small_kernel");
    return RVGP_OK;
}

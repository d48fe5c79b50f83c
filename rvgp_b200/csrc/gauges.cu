// K5 + K6: local-PCA gauges by a one-warp one-sided Jacobi SVD, and the manifold-dimension statistic.
//
// Replaces (reference RVGP/lib/ptu_dijkstra.pyx:396-434): gather the K+1 geodesic neighbours, centre every
// coordinate row by its mean (sequential sum in neighbour order / (K+1)), LAPACK dgesvd('S','N') of the
// D x (K+1) matrix, tangents[i,:,q] = U[:,q], Sigma[i,q] = S[q], failure when S[q] < 1e-10 (pyx:426-428);
// and RVGP/geometry.py:83-97 (manifold_dimension).
//
// One warp per node.  The transposed neighbourhood M^T ((K+1) x D) lives in shared memory; Hestenes
// rotations orthogonalise its D columns (M^T J = Q Sigma), so J holds the LEFT singular vectors of M and the
// column norms are the singular values -- computed to high relative accuracy without forming a covariance.
// Gather-bound: 8 n (K+1) D bytes read, 8 n D (D+1) written.
#include "common.cuh"

namespace rvgp {

constexpr int GAUGE_WARPS = 4;

__global__ void __launch_bounds__(GAUGE_WARPS * 32)
gauges_kernel(const double* __restrict__ X, int n, int D, const int* __restrict__ seq, int Kp1, int dcheck,
              double* __restrict__ tangents, double* __restrict__ Sigma, int* __restrict__ flag) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_warp = Kp1 * D + D * D + 2 * D;
    double* M = smem + warp * per_warp;      // (Kp1, D) row-major: M[r*D + p]
    double* J = M + Kp1 * D;                 // (D, D)  row-major
    double* sig = J + D * D;                 // (D)
    int* rank_of = (int*)(sig + D);          // (D)
    const int i = blockIdx.x * GAUGE_WARPS + warp;
    if (i >= n) return;

    for (int e = lane; e < Kp1 * D; e += 32) {
        const int r = e / D, p = e % D;
        M[e] = __ldg(X + (int64_t)__ldg(seq + (int64_t)i * Kp1 + r) * D + p);
    }
    for (int e = lane; e < D * D; e += 32) J[e] = (e / D == e % D) ? 1.0 : 0.0;
    __syncwarp();
    for (int p = lane; p < D; p += 32) {     // pyx:397-404
        double mean = 0.0;
        for (int r = 0; r < Kp1; ++r) mean += M[r * D + p];
        mean = mean / (double)Kp1;
        for (int r = 0; r < Kp1; ++r) M[r * D + p] -= mean;
    }
    __syncwarp();

    const double eps = 1.0e-15;
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < D - 1; ++p) {
            for (int q = p + 1; q < D; ++q) {
                double a = 0.0, b = 0.0, g = 0.0;
                for (int r = lane; r < Kp1; r += 32) {
                    const double mp = M[r * D + p], mq = M[r * D + q];
                    a = fma(mp, mp, a); b = fma(mq, mq, b); g = fma(mp, mq, g);
                }
                a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
                if (fabs(g) > eps * sqrt(a * b) && g != 0.0) {
                    rotated = true;
                    const double zeta = (b - a) / (2.0 * g);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                    for (int r = lane; r < Kp1; r += 32) {
                        const double mp = M[r * D + p], mq = M[r * D + q];
                        M[r * D + p] = c * mp - s * mq;
                        M[r * D + q] = s * mp + c * mq;
                    }
                    for (int r = lane; r < D; r += 32) {
                        const double jp = J[r * D + p], jq = J[r * D + q];
                        J[r * D + p] = c * jp - s * jq;
                        J[r * D + q] = s * jp + c * jq;
                    }
                    __syncwarp();
                }
            }
        }
        if (!rotated) break;
    }
    // singular values = column norms; order descending (LAPACK convention)
    for (int p = lane; p < D; p += 32) {
        double a = 0.0;
        for (int r = 0; r < Kp1; ++r) a = fma(M[r * D + p], M[r * D + p], a);
        sig[p] = sqrt(a);
    }
    __syncwarp();
    for (int p = lane; p < D; p += 32) {
        int rk = 0;
        for (int q = 0; q < D; ++q) rk += (sig[q] > sig[p]) || (sig[q] == sig[p] && q < p);
        rank_of[p] = rk;
        Sigma[(int64_t)i * D + rk] = sig[p];
        if (rk < dcheck && sig[p] < 1e-10) atomicOr(flag, 1);     // pyx:426-428
    }
    __syncwarp();
    for (int e = lane; e < D * D; e += 32) {
        const int r = e / D, p = e % D;
        tangents[(int64_t)i * D * D + r * D + rank_of[p]] = J[e];
    }
}

// geometry.py:89-91: Sigma**2, normalise each row, cumulative sum along the row
__global__ void cumvar_kernel(const double* __restrict__ Sigma, int n, int D, double* __restrict__ cum) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double tot = 0.0;
    for (int p = 0; p < D; ++p) { const double s = Sigma[i * D + p]; tot += s * s; }
    double run = 0.0;
    for (int p = 0; p < D; ++p) { const double s = Sigma[i * D + p]; run += (s * s) / tot; cum[i * D + p] = run; }
}

// (n, D, dfull) -> (n, D, d): keep the first d columns of every frame (dataclass.py:38)
__global__ void slice_frames_kernel(const double* __restrict__ T, int64_t n, int D, int dfull, int d, double* __restrict__ G) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * D * d) return;
    const int q = (int)(idx % d);
    const int64_t rp = idx / d;
    G[idx] = T[rp * dfull + q];
}

}  // namespace rvgp

using namespace rvgp;

// X (n, D); seq (n, Kp1) neighbourhood indices; tangents (n, D, D) [i][row][q] = q-th left singular vector;
// Sigma (n, D) descending singular values; flag (device int32, zeroed here): bit0 set if any of the first
// `dcheck` singular values of any node is < 1e-10 (-> RVGP_ERR_RANK_DEFICIENT at the caller).
extern "C" int rvgp_tangent_frames(rvgp_handle_t hh, const double* X, int n, int D, const int32_t* seq, int Kp1, int dcheck,
                                   double* tangents, double* Sigma, int32_t* flag) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, n >= 1 && D >= 1 && Kp1 >= 2, "tangent_frames: bad sizes");
    RVGP_REQUIRE(h, D <= dcheck || dcheck <= D, "tangent_frames");
    const size_t smem = (size_t)GAUGE_WARPS * (Kp1 * D + D * D + 2 * D) * sizeof(double);
    RVGP_REQUIRE(h, smem <= 200 * 1024, "tangent_frames: neighbourhood does not fit in shared memory");
    RVGP_CUDA_OK(h, cudaFuncSetAttribute(gauges_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RVGP_CUDA_OK(h, cudaMemsetAsync(flag, 0, sizeof(int), h->stream));
    gauges_kernel<<<cdiv(n, GAUGE_WARPS), GAUGE_WARPS * 32, smem, h->stream>>>(X, n, D, seq, Kp1, dcheck, tangents, Sigma, flag);
    RVGP_LAUNCH_OK(h, "gauges_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_sigma_cumvar(rvgp_handle_t hh, const double* Sigma, int n, int D, double* cum) {
    Handle* h = H(hh);
    cumvar_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(Sigma, n, D, cum);
    RVGP_LAUNCH_OK(h, "cumvar_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_slice_frames(rvgp_handle_t hh, const double* T, int64_t n, int D, int dfull, int d, double* G) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, d >= 1 && d <= dfull, "slice_frames: need 1 <= d <= dfull");
    slice_frames_kernel<<<cdiv(n * D * d, 256), 256, 0, h->stream>>>(T, n, D, dfull, d, G);
    RVGP_LAUNCH_OK(h, "slice_frames_kernel");
    return RVGP_OK;
}

// K17: squared-exponential ("RBF") kernel on feature rows and its reverse-mode adjoint.
//
// Replaces gpflow.kernels.RBF() as reached from train_gp(kernel='rbf') (reference RVGP/main.py:33-37): the channel-wise
// baseline on the scalar-Laplacian eigenvectors.  GPflow evaluates (gpflow/kernels/stationaries.py)
//     r2 = |x/l|^2 + |x'/l|^2 - 2 (x/l).(x'/l),   K = variance * exp(-0.5 * r2),   K_diag = variance
// and differentiates it with TensorFlow's reverse mode.  Here the inner products P = X X'^T come from the FP64 dgemm (K10)
// and these kernels do the HBM-bound elementwise part:
//   rbf_from_dot_kernel   K[i,j] = variance * exp(-0.5 * (xa2[i] + xb2[j] - 2 P[i,j]) / l^2)            (in place over P)
//   rbf_adjoint_kernel    H = Gbar o K (optional output), per-(row, chunk) partial sums of H and H * r2, reduced in a
//                         fixed order by rbf_adjoint_reduce_kernel:  sums[0] = sum H   (= variance * dF/dvariance)
//                                                                    sums[1] = sum H r2 (= l * dF/dl, r2 already / l^2)
//                                                                    rowsum[i] = sum_j H[i,j]  (for dF/dXA)
// One pass over Gbar and P each: 16 m n bytes read (+ 8 m n written when H is wanted); no atomics -> deterministic.
#include "common.cuh"

namespace rvgp {

constexpr int RBF_TX = 256;        // threads per CTA = columns per inner step
constexpr int RBF_CH = 2048;       // columns per CTA chunk

__global__ void __launch_bounds__(RBF_TX)
rbf_from_dot_kernel(int m, int n, const double* P, int64_t ldp, const double* __restrict__ xa2,
                    const double* __restrict__ xb2, double variance, double inv_l2, double* Kout, int64_t ldk) {   // Kout may alias P
    const int i = blockIdx.y;
    const double a = xa2[i];
    const int j0 = blockIdx.x * RBF_CH;
    const int j1 = min(n, j0 + RBF_CH);
    for (int j = j0 + threadIdx.x; j < j1; j += RBF_TX) {
        const double r2 = (a + __ldg(xb2 + j) - 2.0 * P[(int64_t)i * ldp + j]) * inv_l2;
        Kout[(int64_t)i * ldk + j] = variance * exp(-0.5 * r2);
    }
}

__global__ void __launch_bounds__(RBF_TX)
rbf_adjoint_kernel(int m, int n, const double* Gbar, int64_t ldg, const double* P, int64_t ldp,   // Hout may alias either
                   const double* __restrict__ xa2, const double* __restrict__ xb2, double variance, double inv_l2,
                   double* Hout, int64_t ldh, double* __restrict__ part) {
    __shared__ double s0[RBF_TX / 32], s1[RBF_TX / 32];
    const int i = blockIdx.y;
    const double a = xa2[i];
    const int j0 = blockIdx.x * RBF_CH;
    const int j1 = min(n, j0 + RBF_CH);
    double h0 = 0.0, h1 = 0.0;
    for (int j = j0 + threadIdx.x; j < j1; j += RBF_TX) {
        const double r2 = (a + __ldg(xb2 + j) - 2.0 * P[(int64_t)i * ldp + j]) * inv_l2;
        const double hv = Gbar[(int64_t)i * ldg + j] * (variance * exp(-0.5 * r2));
        if (Hout) Hout[(int64_t)i * ldh + j] = hv;
        h0 += hv;
        h1 = fma(hv, r2, h1);
    }
    h0 = warp_sum(h0);
    h1 = warp_sum(h1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s0[warp] = h0; s1[warp] = h1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int w = 0; w < RBF_TX / 32; ++w) { t0 += s0[w]; t1 += s1[w]; }
        const int64_t o = ((int64_t)i * gridDim.x + blockIdx.x) * 2;
        part[o] = t0;
        part[o + 1] = t1;
    }
}

// single CTA, fixed order: rowsum[i] = sum of row i's chunk partials; sums = totals over rows
__global__ void __launch_bounds__(256)
rbf_adjoint_reduce_kernel(int m, int nchunk, const double* __restrict__ part, double* __restrict__ rowsum, double* __restrict__ sums) {
    __shared__ double s0[256], s1[256];
    double t0 = 0.0, t1 = 0.0;
    for (int i = threadIdx.x; i < m; i += 256) {
        double r0 = 0.0, r1 = 0.0;
        for (int c = 0; c < nchunk; ++c) { r0 += part[((int64_t)i * nchunk + c) * 2]; r1 += part[((int64_t)i * nchunk + c) * 2 + 1]; }
        if (rowsum) rowsum[i] = r0;
        t0 += r0;
        t1 += r1;
    }
    s0[threadIdx.x] = t0;
    s1[threadIdx.x] = t1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s0[threadIdx.x] += s0[threadIdx.x + o]; s1[threadIdx.x] += s1[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { sums[0] = s0[0]; sums[1] = s1[0]; }
}

// out[i, :] = (HX[i, :] - rowsum[i] * XA[i, :]) * inv_l2      dF/dXA of the RBF kernel (m x k, small)
__global__ void rbf_dx_kernel(int64_t m, int k, const double* __restrict__ HX, int64_t ldhx, const double* __restrict__ rowsum,
                              const double* __restrict__ XA, int64_t ldx, double inv_l2, double* __restrict__ out, int64_t ldo) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * k) return;
    const int64_t i = idx / k;
    const int c = (int)(idx % k);
    out[i * ldo + c] = (HX[i * ldhx + c] - rowsum[i] * XA[i * ldx + c]) * inv_l2;
}

// A[r, c] = alpha * A[r, c] + beta * (r == c)      (scaled copy-in-place + diagonal shift: B = I + A A^T / s2, I - B^-1, ...)
__global__ void scale_shift_kernel(int64_t nrows, int ncols, double alpha, double beta, double* __restrict__ A, int64_t lda) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrows * ncols) return;
    const int64_t r = idx / ncols;
    const int c = (int)(idx % ncols);
    A[r * lda + c] = alpha * A[r * lda + c] + (r == c ? beta : 0.0);
}

}  // namespace rvgp

using namespace rvgp;

extern "C" int rvgp_rbf_from_dot_f64(rvgp_handle_t hh, int m, int n, const double* P, int64_t ldp, const double* xa2,
                                     const double* xb2, double variance, double lengthscale, double* Kout, int64_t ldk) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, m >= 0 && n >= 0 && ldp >= n && ldk >= n && lengthscale > 0.0, "rbf_from_dot: bad sizes");
    if (m == 0 || n == 0) return RVGP_OK;
    const double inv_l2 = 1.0 / (lengthscale * lengthscale);
    for (int r0 = 0; r0 < m; r0 += 65535) {          // gridDim.y limit
        const int mr = (m - r0 < 65535) ? m - r0 : 65535;
        rbf_from_dot_kernel<<<dim3(cdiv(n, RBF_CH), mr), RBF_TX, 0, h->stream>>>(mr, n, P + (int64_t)r0 * ldp, ldp, xa2 + r0, xb2,
                                                                                 variance, inv_l2, Kout + (int64_t)r0 * ldk, ldk);
        RVGP_LAUNCH_OK(h, "rbf_from_dot_kernel");
    }
    return RVGP_OK;
}

extern "C" int64_t rvgp_rbf_adjoint_workspace_bytes(int m, int n) {
    return (int64_t)m * cdiv(n > 0 ? n : 1, RBF_CH) * 2 * 8 + 16;
}

extern "C" int rvgp_rbf_adjoint_f64(rvgp_handle_t hh, int m, int n, const double* Gbar, int64_t ldg, const double* P, int64_t ldp,
                                    const double* xa2, const double* xb2, double variance, double lengthscale, double* Hout,
                                    int64_t ldh, double* rowsum, double* sums, void* workspace, int64_t workspace_bytes) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, m >= 1 && n >= 1 && ldg >= n && ldp >= n && lengthscale > 0.0 && sums != nullptr, "rbf_adjoint: bad args");
    RVGP_REQUIRE(h, Hout == nullptr || ldh >= n, "rbf_adjoint: bad ldh");
    if (rvgp_rbf_adjoint_workspace_bytes(m, n) > workspace_bytes)
        return set_error(h, RVGP_ERR_CAPACITY, "rbf_adjoint: workspace too small%s%s");
    const double inv_l2 = 1.0 / (lengthscale * lengthscale);
    const int nchunk = cdiv(n, RBF_CH);
    double* part = (double*)workspace;
    for (int r0 = 0; r0 < m; r0 += 65535) {
        const int mr = (m - r0 < 65535) ? m - r0 : 65535;
        rbf_adjoint_kernel<<<dim3(nchunk, mr), RBF_TX, 0, h->stream>>>(
            mr, n, Gbar + (int64_t)r0 * ldg, ldg, P + (int64_t)r0 * ldp, ldp, xa2 + r0, xb2, variance, inv_l2,
            Hout ? Hout + (int64_t)r0 * ldh : nullptr, ldh, part + (int64_t)r0 * nchunk * 2);
        RVGP_LAUNCH_OK(h, "rbf_adjoint_kernel");
    }
    rbf_adjoint_reduce_kernel<<<1, 256, 0, h->stream>>>(m, nchunk, part, rowsum, sums);
    RVGP_LAUNCH_OK(h, "rbf_adjoint_reduce_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_rbf_dx_f64(rvgp_handle_t hh, int64_t m, int k, const double* HX, int64_t ldhx, const double* rowsum,
                               const double* XA, int64_t ldx, double lengthscale, double* out, int64_t ldo) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, m >= 0 && k >= 0 && lengthscale > 0.0, "rbf_dx: bad args");
    if (m * k == 0) return RVGP_OK;
    rbf_dx_kernel<<<cdiv(m * k, 256), 256, 0, h->stream>>>(m, k, HX, ldhx, rowsum, XA, ldx, 1.0 / (lengthscale * lengthscale), out, ldo);
    RVGP_LAUNCH_OK(h, "rbf_dx_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_scale_shift_f64(rvgp_handle_t hh, int64_t nrows, int ncols, double alpha, double beta, double* A, int64_t lda) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, nrows >= 0 && ncols >= 0 && lda >= ncols, "scale_shift: bad args");
    if (nrows * ncols == 0) return RVGP_OK;
    scale_shift_kernel<<<cdiv(nrows * ncols, 256), 256, 0, h->stream>>>(nrows, ncols, alpha, beta, A, lda);
    RVGP_LAUNCH_OK(h, "scale_shift_kernel");
    return RVGP_OK;
}

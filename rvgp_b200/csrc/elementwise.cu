// K12 helpers: small HBM-bound elementwise / row-wise kernels used by vector-diffusion smoothing
// (reference RVGP/smoothing.py:37-64) and random_vector_field (RVGP/dataclass.py:91-103).
#include "common.cuh"

namespace rvgp {

__global__ void axpy_kernel(int64_t nrows, int ncols, double a, const double* __restrict__ X, int64_t ldx,
                            double* __restrict__ Y, int64_t ldy) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrows * ncols) return;
    const int64_t r = idx / ncols;
    const int c = (int)(idx % ncols);
    Y[r * ldy + c] = fma(a, __ldg(X + r * ldx + c), Y[r * ldy + c]);
}

// out[i] = ||x[i, :]||_2 (sum of squares in column order, like np.linalg.norm(axis=-1))
__global__ void row_norm_kernel(int64_t n, int d, const double* __restrict__ x, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < d; ++k) { const double v = x[i * d + k]; s = fma(v, v, s); }
    out[i] = sqrt(s);
}

// smoothing.py:62   out = out * out_abs / (ind * ||out||)     (ind, out_abs nullable -> 1)
__global__ void renorm_rows_kernel(int64_t n, int d, double* __restrict__ x, const double* __restrict__ out_abs,
                                   const double* __restrict__ ind) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < d; ++k) { const double v = x[i * d + k]; s = fma(v, v, s); }
    const double num = out_abs ? out_abs[i] : 1.0;
    const double den = (ind ? ind[i] : 1.0) * sqrt(s);
    for (int k = 0; k < d; ++k) x[i * d + k] = x[i * d + k] * num / den;
}

}  // namespace rvgp

using namespace rvgp;

extern "C" int rvgp_axpy_f64(rvgp_handle_t hh, int64_t nrows, int ncols, double a, const double* X, int64_t ldx, double* Y,
                             int64_t ldy) {
    Handle* h = H(hh);
    if (nrows * ncols == 0) return RVGP_OK;
    axpy_kernel<<<cdiv(nrows * ncols, 256), 256, 0, h->stream>>>(nrows, ncols, a, X, ldx, Y, ldy);
    RVGP_LAUNCH_OK(h, "axpy_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_row_norms_f64(rvgp_handle_t hh, int64_t n, int d, const double* x, double* out) {
    Handle* h = H(hh);
    if (n == 0) return RVGP_OK;
    row_norm_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(n, d, x, out);
    RVGP_LAUNCH_OK(h, "row_norm_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_renorm_rows_f64(rvgp_handle_t hh, int64_t n, int d, double* x, const double* out_abs, const double* ind) {
    Handle* h = H(hh);
    if (n == 0) return RVGP_OK;
    renorm_rows_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(n, d, x, out_abs, ind);
    RVGP_LAUNCH_OK(h, "renorm_rows_kernel");
    return RVGP_OK;
}

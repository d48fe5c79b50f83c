// K1: furthest-point sampling as ONE persistent cooperative kernel (no N x N matrix).
//
// Replaces furthest_point_sampling (reference RVGP/geometry.py:126-162), which forms the full
// sklearn.metrics.pairwise_distances(x) matrix (geometry.py:142) and then loops
//     idx = argmax(ds) (first max = LOWEST index); lambdas[i] = ds[idx]; ds = min(ds, D[idx]).
// Distances use sklearn's float64 euclidean_distances expansion so values match the reference:
//     D[i,j] = sqrt(max((-2 x_i.x_j + |x_i|^2) + |x_j|^2, 0)),  D[i,i] = 0.
// Per sample step every thread updates its slice of ds, the argmax is reduced by warp shuffles
// (ties -> lowest index), block results are exchanged through global memory and ONE grid-wide barrier per step
// (double-buffered).  Latency / L2-bound: 8 n (D+2) bytes per step, sequential over samples.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace rvgp {

__device__ __forceinline__ double sk_dist2(const double* __restrict__ X, const double* __restrict__ xx, int D,
                                           const double* __restrict__ xi, double xxi, int64_t j) {
    double dot = 0.0;
    for (int k = 0; k < D; ++k) dot = fma(xi[k], __ldg(X + j * D + k), dot);
    double d = -2.0 * dot;
    d = __dadd_rn(d, xxi);
    d = __dadd_rn(d, __ldg(xx + j));
    return d > 0.0 ? d : 0.0;
}

__global__ void row_norms_kernel(const double* __restrict__ X, int n, int D, double* __restrict__ xx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < D; ++k) { const double v = X[i * D + k]; s = __dadd_rn(s, __dmul_rn(v, v)); }
    xx[i] = s;
}

struct Best { double v; int i; };
__device__ __forceinline__ Best better(Best a, Best b) {   // larger value wins; ties -> lower index
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

constexpr int FPS_MAXD = 1024;       // feature-space FPS for SGPR inducing points: D = n_eigenpairs (main.py:60)
constexpr int FPS_THREAD_MAXD = 64;  // above this one WARP (not one thread) owns a point: coalesced row reads

__global__ void __launch_bounds__(256)
fps_kernel(const double* __restrict__ X, const double* __restrict__ xx, int n, int D, int n_out, int use_spacing,
           double spacing, const double* __restrict__ diam_p, int start_idx, int* __restrict__ perm,
           double* __restrict__ lambdas, int* __restrict__ count, double* __restrict__ ds, double* __restrict__ blk_v,
           int* __restrict__ blk_i) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double xi[FPS_MAXD];
    __shared__ Best wbest[8];
    __shared__ Best gbest;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double diam = use_spacing ? *diam_p : 1.0;
    int cur = start_idx;
    if (tid == 0) { perm[0] = start_idx; lambdas[0] = 0.0; }
    int produced = 1;
    for (int step = 1; step < n_out; ++step) {
        // row `cur` of the distance matrix -> ds = min(ds, D[cur]) (step 1: ds = D[start])
        for (int k = threadIdx.x; k < D; k += blockDim.x) xi[k] = X[(int64_t)cur * D + k];
        __syncthreads();
        const double xxi = xx[cur];
        Best b{-1.0, 0x7fffffff};
        if (D <= FPS_THREAD_MAXD) {
            for (int j = tid; j < n; j += nth) {
                double d = (j == cur) ? 0.0 : sqrt(sk_dist2(X, xx, D, xi, xxi, j));
                if (step > 1) d = fmin(ds[j], d);
                ds[j] = d;
                if (d > b.v) { b.v = d; b.i = j; }        // ascending j: strict > keeps the first maximum
            }
        } else {
            // high-dimensional rows (feature space): lanes stride over the coordinates, every lane ends with the same d
            for (int j = tid >> 5; j < n; j += nth >> 5) {
                double dot = 0.0;
                for (int k = lane; k < D; k += 32) dot = fma(xi[k], __ldg(X + (int64_t)j * D + k), dot);
                dot = warp_sum(dot);
                double d = -2.0 * dot;
                d = __dadd_rn(d, xxi);
                d = __dadd_rn(d, __ldg(xx + j));
                d = (j == cur) ? 0.0 : sqrt(d > 0.0 ? d : 0.0);
                if (step > 1) d = fmin(ds[j], d);
                if (lane == 0) ds[j] = d;
                if (d > b.v) { b.v = d; b.i = j; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Best c{__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.i, o)};
            b = better(b, c);
        }
        if (lane == 0) wbest[warp] = b;
        __syncthreads();
        const int buf = (step & 1) * gridDim.x;
        if (threadIdx.x == 0) {
            Best r = wbest[0];
            for (int w = 1; w < 8; ++w) r = better(r, wbest[w]);
            blk_v[buf + blockIdx.x] = r.v;
            blk_i[buf + blockIdx.x] = r.i;
        }
        grid.sync();
        // every block reduces all block results redundantly (no second barrier needed)
        if (warp == 0) {
            Best r{-1.0, 0x7fffffff};
            for (int k = lane; k < (int)gridDim.x; k += 32) r = better(r, Best{blk_v[buf + k], blk_i[buf + k]});
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                Best c{__shfl_xor_sync(0xffffffffu, r.v, o), __shfl_xor_sync(0xffffffffu, r.i, o)};
                r = better(r, c);
            }
            if (lane == 0) gbest = r;
        }
        __syncthreads();
        const Best g = gbest;
        __syncthreads();
        if (use_spacing && g.v / diam < spacing) break;      // geometry.py:156-160: truncate to [:step]
        if (tid == 0) { perm[step] = g.i; lambdas[step] = g.v; }
        produced = step + 1;
        cur = g.i;
    }
    if (tid == 0) *count = produced;
}

// all-pairs maximum of sklearn's squared distance expansion (diam = D.max(), geometry.py:144)
__global__ void __launch_bounds__(256)
pairmax_kernel(const double* __restrict__ X, const double* __restrict__ xx, int n, int D, unsigned long long* __restrict__ out) {
    __shared__ double tile[256 * 8];
    __shared__ double txx[256];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double xi[8];
    double best = 0.0;
    // D <= 8 fast path keeps the point in registers; larger D streams it from global memory
    const bool small = D <= 8;
    if (small)
        for (int k = 0; k < 8; ++k) xi[k] = (i < n && k < D) ? X[(int64_t)i * D + k] : 0.0;
    const double xxi = (i < n) ? xx[i] : 0.0;
    for (int j0 = 0; j0 < n; j0 += 256) {
        const int tl = min(256, n - j0);
        __syncthreads();
        if (small) {
            for (int e = threadIdx.x; e < tl * D; e += 256) tile[(e / D) * 8 + (e % D)] = X[(int64_t)j0 * D + e];
        }
        if (threadIdx.x < tl) txx[threadIdx.x] = xx[j0 + threadIdx.x];
        __syncthreads();
        if (i >= n) continue;
        for (int c = 0; c < tl; ++c) {
            double dot = 0.0;
            if (small) {
                for (int k = 0; k < D; ++k) dot = fma(xi[k], tile[c * 8 + k], dot);
            } else {
                for (int k = 0; k < D; ++k) dot = fma(__ldg(X + (int64_t)i * D + k), __ldg(X + (int64_t)(j0 + c) * D + k), dot);
            }
            double d = -2.0 * dot;
            d = __dadd_rn(d, xxi);
            d = __dadd_rn(d, txx[c]);
            if (j0 + c != i) best = fmax(best, d);
        }
    }
    best = fmax(best, 0.0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(best));   // best >= 0: bit order == value order
}

__global__ void sqrt_bits_kernel(const unsigned long long* __restrict__ in, double* __restrict__ out) {
    *out = sqrt(__longlong_as_double((long long)*in));
}

}  // namespace rvgp

using namespace rvgp;

extern "C" int64_t rvgp_fps_workspace_bytes(rvgp_handle_t hh, int n) {
    const int64_t blocks = H(hh)->sm_count;
    return (int64_t)n * 8 * 2 + blocks * 2 * (8 + 4) + 64;
}

// X (n, D) FP64.  N > 0: exactly N samples (geometry.py: N given); N == 0: run until lambdas[i]/diam < spacing
// (diam = max pairwise distance, computed here).  perm / lambdas need capacity N (or n when N == 0);
// count_out (device int32) = number of valid entries.
extern "C" int rvgp_fps_f64(rvgp_handle_t hh, const double* X, int n, int D, int N, double spacing, int start_idx,
                            int32_t* perm, double* lambdas, int32_t* count_out, void* workspace, int64_t workspace_bytes) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, n >= 1 && D >= 1 && D <= FPS_MAXD, "fps: D must be in [1,1024]");
    RVGP_REQUIRE(h, start_idx >= 0 && start_idx < n && N >= 0 && N <= n, "fps: bad start_idx / N");
    if (rvgp_fps_workspace_bytes(hh, n) > workspace_bytes) return set_error(h, RVGP_ERR_CAPACITY, "fps: workspace too small%s%s");
    int blocks = h->sm_count;
    char* p = (char*)workspace;
    double* xx = (double*)p; p += (int64_t)n * 8;
    double* ds = (double*)p; p += (int64_t)n * 8;
    double* blk_v = (double*)p; p += (int64_t)blocks * 2 * 8;
    int* blk_i = (int*)p; p += (int64_t)blocks * 2 * 4;
    p = (char*)(((uintptr_t)p + 15) / 16 * 16);
    unsigned long long* dbits = (unsigned long long*)p; p += 8;
    double* diam = (double*)p;
    row_norms_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(X, n, D, xx);
    RVGP_LAUNCH_OK(h, "row_norms_kernel");
    const int use_spacing = (N == 0);
    if (use_spacing) {
        RVGP_CUDA_OK(h, cudaMemsetAsync(dbits, 0, 8, h->stream));
        pairmax_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(X, xx, n, D, dbits);
        RVGP_LAUNCH_OK(h, "pairmax_kernel");
        sqrt_bits_kernel<<<1, 1, 0, h->stream>>>(dbits, diam);
        RVGP_LAUNCH_OK(h, "sqrt_bits_kernel");
    }
    int n_out = use_spacing ? n : N;
    void* args[] = {(void*)&X, (void*)&xx, (void*)&n, (void*)&D, (void*)&n_out, (void*)&use_spacing, (void*)&spacing,
                    (void*)&diam, (void*)&start_idx, (void*)&perm, (void*)&lambdas, (void*)&count_out, (void*)&ds,
                    (void*)&blk_v, (void*)&blk_i};
    RVGP_CUDA_OK(h, cudaLaunchCooperativeKernel((void*)fps_kernel, dim3(blocks), dim3(256), args, 0, h->stream));
    h->launches++;
    return RVGP_OK;
}

// K3: kNN lists -> symmetric CSR with self loops (int32, sorted columns), plus the locality
// (Morton) ordering and CSR permutation used by the eigensolver.
//
// Replaces: `A += eye; nx.from_scipy_sparse_array(A)` = UNION symmetrisation (reference
// RVGP/geometry.py:111-112) and the CSR rebuild `csgraph.maximum(csgraph.T)` seen by the Cython code
// (RVGP/lib/ptu_dijkstra.pyx:84-103): symmetric, sorted column indices, diagonal present.
// HBM-bound integer work: 64-bit (row,col) keys, CUB radix sort + unique, row boundaries by comparison.
#include <cub/cub.cuh>

#include "common.cuh"

namespace rvgp {

__global__ void make_edge_keys_kernel(const int* __restrict__ knn, int n, int k, unsigned long long* __restrict__ keys) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nk = (int64_t)n * k;
    if (idx < nk) {
        const unsigned long long i = (unsigned long long)(idx / k);
        const unsigned long long j = (unsigned long long)__ldg(knn + idx);
        keys[2 * idx] = (i << 32) | j;
        keys[2 * idx + 1] = (j << 32) | i;
    } else if (idx < nk + n) {
        const unsigned long long i = (unsigned long long)(idx - nk);
        keys[2 * nk + i] = (i << 32) | i;   // self loop (A += eye)
    }
}

__global__ void keys_to_csr_kernel(const unsigned long long* __restrict__ ukeys, const int* __restrict__ nnz_p, int n,
                                   int* __restrict__ indptr, int* __restrict__ indices) {
    const int nnz = *nnz_p;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e == 0) indptr[n] = nnz;
    if (e >= nnz) return;
    const unsigned long long key = ukeys[e];
    const int row = (int)(key >> 32);
    indices[e] = (int)(key & 0xffffffffull);
    if (e == 0 || (int)(ukeys[e - 1] >> 32) != row) indptr[row] = (int)e;   // every row has its self loop
}

struct CsrWs {
    unsigned long long *keys, *sorted, *uniq;
    int* nsel;
    void* cub;
    size_t cub_bytes, total;
};

static CsrWs csr_ws_layout(int n, int k, void* base) {
    const size_t E = (size_t)2 * n * k + n;
    size_t sort_b = 0, uniq_b = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, sort_b, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)E);
    cub::DeviceSelect::Unique(nullptr, uniq_b, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr, (int)E);
    CsrWs w;
    char* p = (char*)base;
    auto take = [&](size_t b) { char* r = p; p += (b + 255) / 256 * 256; return r; };
    w.keys = (unsigned long long*)take(E * 8);
    w.sorted = (unsigned long long*)take(E * 8);
    w.uniq = (unsigned long long*)take(E * 8);
    w.nsel = (int*)take(256);
    w.cub_bytes = sort_b > uniq_b ? sort_b : uniq_b;
    w.cub = take(w.cub_bytes);
    w.total = (size_t)(p - (char*)base);
    return w;
}

// ---- locality ordering --------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long enc_double(double v) {   // order-preserving double -> u64
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_double(unsigned long long u) {
    u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)u);
}

__global__ void bbox_kernel(const double* __restrict__ X, int n, int D, int nd, unsigned long long* __restrict__ mm) {
    // mm[2*j] = min over points of coordinate j (encoded), mm[2*j+1] = max
    for (int j = 0; j < nd; ++j) {
        double lo = __longlong_as_double(0x7ff0000000000000ll), hi = -lo;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            const double v = __ldg(X + i * D + j);
            lo = fmin(lo, v); hi = fmax(hi, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(mm + 2 * j, enc_double(lo));
            atomicMax(mm + 2 * j + 1, enc_double(hi));
        }
    }
}

__device__ __forceinline__ unsigned long long spread3(unsigned long long v) {   // 21 bits -> every third bit
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__global__ void morton_kernel(const double* __restrict__ X, int n, int D, int nd, const unsigned long long* __restrict__ mm,
                              unsigned long long* __restrict__ codes, int* __restrict__ ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long code = 0;
    for (int j = 0; j < nd; ++j) {
        const double lo = dec_double(mm[2 * j]), hi = dec_double(mm[2 * j + 1]);
        const double w = hi - lo;
        double u = (w > 0) ? (__ldg(X + i * D + j) - lo) / w : 0.0;
        u = fmin(fmax(u, 0.0), 1.0);
        const unsigned long long q = (unsigned long long)(u * 2097151.0);
        code |= spread3(q) << j;
    }
    codes[i] = code;
    ids[i] = (int)i;
}

__global__ void invert_perm_kernel(const int* __restrict__ order, int n, int* __restrict__ inv) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) inv[order[r]] = (int)r;
}

// ---- CSR permutation ------------------------------------------------------------------------------------
__global__ void perm_rowlen_kernel(const int* __restrict__ indptr, const int* __restrict__ order, int n, int* __restrict__ len) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) { const int o = order[r]; len[r] = indptr[o + 1] - indptr[o]; }
    if (r == n) len[n] = 0;
}

__global__ void perm_fill_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, const int* __restrict__ order,
                                 const int* __restrict__ inv, int n, const int* __restrict__ new_indptr,
                                 int* __restrict__ new_indices) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int o = order[r];
    const int s = indptr[o], len = indptr[o + 1] - s, dst = new_indptr[r];
    // relabel + insertion sort (rows hold ~10-40 entries)
    for (int e = 0; e < len; ++e) {
        const int v = inv[indices[s + e]];
        int p = e;
        while (p > 0 && new_indices[dst + p - 1] > v) { new_indices[dst + p] = new_indices[dst + p - 1]; --p; }
        new_indices[dst + p] = v;
    }
}

}  // namespace rvgp

using namespace rvgp;

extern "C" int64_t rvgp_knn_to_csr_workspace_bytes(int n, int k) { return (int64_t)csr_ws_layout(n, k, nullptr).total; }

// knn: (n, k) int32 directed neighbour lists (self excluded).  Outputs: indptr (n+1), indices (capacity
// 2*n*k + n), nnz_out (device int32).  Asynchronous; read *nnz_out after synchronising.
extern "C" int rvgp_knn_to_csr(rvgp_handle_t hh, const int32_t* knn, int n, int k, int32_t* indptr, int32_t* indices,
                               int32_t* nnz_out, void* workspace, int64_t workspace_bytes) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, n >= 1 && k >= 1, "knn_to_csr: bad sizes");
    RVGP_REQUIRE(h, (int64_t)2 * n * k + n < 2147483647ll, "knn_to_csr: more than 2^31 edges");
    CsrWs w = csr_ws_layout(n, k, workspace);
    if ((int64_t)w.total > workspace_bytes) return set_error(h, RVGP_ERR_CAPACITY, "knn_to_csr: workspace too small%s%s");
    const int64_t E = (int64_t)2 * n * k + n;
    make_edge_keys_kernel<<<cdiv((int64_t)n * k + n, 256), 256, 0, h->stream>>>(knn, n, k, w.keys);
    RVGP_LAUNCH_OK(h, "make_edge_keys_kernel");
    int bits = 1;
    while ((1ll << bits) < n) ++bits;
    size_t cb = w.cub_bytes;
    RVGP_CUDA_OK(h, cub::DeviceRadixSort::SortKeys(w.cub, cb, w.keys, w.sorted, (int)E, 0, 32 + bits, h->stream));
    cb = w.cub_bytes;
    RVGP_CUDA_OK(h, cub::DeviceSelect::Unique(w.cub, cb, w.sorted, w.uniq, w.nsel, (int)E, h->stream));
    h->launches += 4;
    keys_to_csr_kernel<<<cdiv(E, 256), 256, 0, h->stream>>>(w.uniq, w.nsel, n, indptr, indices);
    RVGP_LAUNCH_OK(h, "keys_to_csr_kernel");
    RVGP_CUDA_OK(h, cudaMemcpyAsync(nnz_out, w.nsel, sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    return RVGP_OK;
}

static size_t morton_ws(int n, size_t* cub_bytes) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr,
                                    (int*)nullptr, n);
    if (cub_bytes) *cub_bytes = b;
    const size_t a = ((size_t)n * 8 + 255) / 256 * 256, c = ((size_t)n * 4 + 255) / 256 * 256;
    return 2 * a + c + 256 + (b + 255) / 256 * 256;
}

extern "C" int64_t rvgp_morton_order_workspace_bytes(int n) { return (int64_t)morton_ws(n, nullptr); }

// Locality-preserving ordering of the points (Morton code of the first min(D,3) coordinates).
// order[r] = original index of the point placed at position r; inv[order[r]] = r.
extern "C" int rvgp_morton_order(rvgp_handle_t hh, const double* X, int n, int D, int32_t* order, int32_t* inv,
                                 void* workspace, int64_t workspace_bytes) {
    Handle* h = H(hh);
    size_t cub_b = 0;
    const size_t need = morton_ws(n, &cub_b);
    if ((int64_t)need > workspace_bytes) return set_error(h, RVGP_ERR_CAPACITY, "morton_order: workspace too small%s%s");
    char* p = (char*)workspace;
    const size_t a = ((size_t)n * 8 + 255) / 256 * 256, c = ((size_t)n * 4 + 255) / 256 * 256;
    unsigned long long* codes = (unsigned long long*)p; p += a;
    unsigned long long* codes2 = (unsigned long long*)p; p += a;
    int* ids = (int*)p; p += c;
    unsigned long long* mm = (unsigned long long*)p; p += 256;
    void* cubtmp = p;
    const int nd = D < 3 ? D : 3;
    unsigned long long init[6] = {~0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull};
    RVGP_CUDA_OK(h, cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    bbox_kernel<<<h->sm_count * 4, 256, 0, h->stream>>>(X, n, D, nd, mm);
    RVGP_LAUNCH_OK(h, "bbox_kernel");
    morton_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(X, n, D, nd, mm, codes, ids);
    RVGP_LAUNCH_OK(h, "morton_kernel");
    RVGP_CUDA_OK(h, cub::DeviceRadixSort::SortPairs(cubtmp, cub_b, codes, codes2, ids, order, n, 0, 63, h->stream));
    h->launches += 3;
    invert_perm_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(order, n, inv);
    RVGP_LAUNCH_OK(h, "invert_perm_kernel");
    return RVGP_OK;
}

static size_t permute_ws(int n, size_t* cub_bytes) {
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (int*)nullptr, (int*)nullptr, n + 1);
    if (cub_bytes) *cub_bytes = b;
    return ((size_t)(n + 1) * 4 + 255) / 256 * 256 + (b + 255) / 256 * 256;
}

extern "C" int64_t rvgp_csr_permute_workspace_bytes(int n) { return (int64_t)permute_ws(n, nullptr); }

// Symmetric permutation of a CSR pattern: new row r = old row order[r], columns relabelled through inv and
// re-sorted.  new_indices has the same length as indices.
extern "C" int rvgp_csr_permute(rvgp_handle_t hh, int n, const int32_t* indptr, const int32_t* indices, const int32_t* order,
                                const int32_t* inv, int32_t* new_indptr, int32_t* new_indices, void* workspace,
                                int64_t workspace_bytes) {
    Handle* h = H(hh);
    size_t cub_b = 0;
    const size_t need = permute_ws(n, &cub_b);
    if ((int64_t)need > workspace_bytes) return set_error(h, RVGP_ERR_CAPACITY, "csr_permute: workspace too small%s%s");
    int* len = (int*)workspace;
    void* cubtmp = (char*)workspace + ((size_t)(n + 1) * 4 + 255) / 256 * 256;
    perm_rowlen_kernel<<<cdiv(n + 1, 256), 256, 0, h->stream>>>(indptr, order, n, len);
    RVGP_LAUNCH_OK(h, "perm_rowlen_kernel");
    RVGP_CUDA_OK(h, cub::DeviceScan::ExclusiveSum(cubtmp, cub_b, len, new_indptr, n + 1, h->stream));
    h->launches += 1;
    perm_fill_kernel<<<cdiv(n, 128), 128, 0, h->stream>>>(indptr, indices, order, inv, n, new_indptr, new_indices);
    RVGP_LAUNCH_OK(h, "perm_fill_kernel");
    return RVGP_OK;
}

// ---- typ='affinity' graph (geometry.py:114-118): dense Gaussian-kernel weights --------------------------------------------
//   A[i][j] = exp(-dist(i,j)^2 / (2 sigma^2)),  dist = sklearn.metrics.pairwise_distances(X) (euclidean_distances expansion:
//   sqrt(max(-2 x.y + |x|^2 + |y|^2, 0)), diagonal forced to 0), squared again like the reference's `** 2`.
// One thread per (i, j); the row of X for i is staged in shared memory.  O(n^2 D) by definition of this graph type.
namespace rvgp {
__global__ void affinity_norms_kernel(const double* __restrict__ X, int n, int D, double* __restrict__ xx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < D; ++k) { const double v = X[i * D + k]; s = __dadd_rn(s, __dmul_rn(v, v)); }
    xx[i] = s;
}

__global__ void __launch_bounds__(256)
affinity_kernel(const double* __restrict__ X, const double* __restrict__ xx, int n, int D, double two_sigma2,
                double* __restrict__ A) {
    extern __shared__ double xi[];
    const int i = blockIdx.y;
    for (int k = threadIdx.x; k < D; k += blockDim.x) xi[k] = X[(int64_t)i * D + k];
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double dot = 0.0;
    for (int k = 0; k < D; ++k) dot = fma(xi[k], __ldg(X + (int64_t)j * D + k), dot);
    double d2 = __dadd_rn(__dadd_rn(-2.0 * dot, xx[i]), xx[j]);
    d2 = d2 > 0.0 ? d2 : 0.0;
    double dist = (i == j) ? 0.0 : sqrt(d2);
    A[(int64_t)i * n + j] = exp(-__dmul_rn(dist, dist) / two_sigma2);
}
}  // namespace rvgp

// Dense affinity matrix A (n x n, row-major) of the reference's typ='affinity' graph; workspace: n doubles.
extern "C" int rvgp_affinity_f64(rvgp_handle_t hh, const double* X, int n, int D, double sigma, double* A, double* workspace) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, n >= 0 && D >= 1 && D <= 4096, "affinity: need 1 <= D <= 4096");
    RVGP_REQUIRE(h, sigma > 0.0, "affinity: sigma must be positive");
    if (n == 0) return RVGP_OK;
    affinity_norms_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(X, n, D, workspace);
    RVGP_LAUNCH_OK(h, "affinity_norms_kernel");
    dim3 grid(cdiv(n, 256), n);
    RVGP_REQUIRE(h, n <= 65535, "affinity: n must be <= 65535 (dense n x n graph)");
    affinity_kernel<<<grid, 256, (size_t)D * sizeof(double), h->stream>>>(X, workspace, n, D, 2.0 * sigma * sigma, A);
    RVGP_LAUNCH_OK(h, "affinity_kernel");
    return RVGP_OK;
}

// K2 (grid-accelerated): exact kNN for D <= 3 through a uniform cell grid.
//
// Same contract and the same arithmetic as knn.cu (squared distance accumulated coordinate by coordinate with
// __dsub_rn/__dmul_rn/__dadd_rn, sample removed by index, ties by lowest index), so the result is bit-identical to the
// brute-force kernel -- only the candidate set is pruned: a query scans the cells of growing cubic shells around its own
// cell and stops after shell r as soon as its k-th best squared distance is strictly below (r*h)^2, the smallest
// possible squared distance of any point outside the scanned cube.  n^2 -> ~n * (points in a few cells).
#include <cub/cub.cuh>

#include "common.cuh"

namespace rvgp {

struct GridParams {
    double lo[3], inv_h, h;
    int g[3];
    int D;
};

__device__ __forceinline__ int cell_coord(double x, double lo, double inv_h, int gmax) {
    int c = (int)floor((x - lo) * inv_h);
    return c < 0 ? 0 : (c >= gmax ? gmax - 1 : c);
}

__global__ void grid_cell_kernel(const double* __restrict__ X, int n, GridParams P, int* __restrict__ cell, int* __restrict__ ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c[3] = {0, 0, 0};
    for (int j = 0; j < P.D; ++j) c[j] = cell_coord(X[(int64_t)i * P.D + j], P.lo[j], P.inv_h, P.g[j]);
    cell[i] = (c[2] * P.g[1] + c[1]) * P.g[0] + c[0];
    ids[i] = i;
}

__global__ void grid_bounds_kernel(const int* __restrict__ cell_sorted, int n, int* __restrict__ cstart, int* __restrict__ cend) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cell_sorted[i];
    if (i == 0 || cell_sorted[i - 1] != c) cstart[c] = i;
    if (i == n - 1 || cell_sorted[i + 1] != c) cend[c] = i + 1;
}

__global__ void grid_gather_kernel(const double* __restrict__ X, const int* __restrict__ ids_sorted, int n, int D,
                                   double* __restrict__ Xs) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * D) return;
    Xs[idx] = X[(int64_t)ids_sorted[idx / D] * D + idx % D];
}

template <int D, int KMAX>
__global__ void __launch_bounds__(128)
knn_grid_kernel(const double* __restrict__ Xs, const int* __restrict__ ids_sorted, const int* __restrict__ cstart,
                const int* __restrict__ cend, int n, GridParams P, int q_begin, int q_count, int k, const int* __restrict__ qpos,
                int* __restrict__ out_idx, double* __restrict__ out_d2) {
    // thread t handles the t-th query of this rank IN CELL-SORTED ORDER (qpos lists the sorted positions of the queries)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= q_count) return;
    const int sp = qpos[t];                      // position in the sorted arrays
    const int qi = ids_sorted[sp];               // original index of the query
    double q[D];
    int qc[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < D; ++j) { q[j] = Xs[(int64_t)sp * D + j]; qc[j] = cell_coord(q[j], P.lo[j], P.inv_h, P.g[j]); }
    double bd[KMAX];
    int bi[KMAX];
#pragma unroll
    for (int r = 0; r < KMAX; ++r) { bd[r] = __longlong_as_double(0x7ff0000000000000ll); bi[r] = 0x7fffffff; }
    double thr = __longlong_as_double(0x7ff0000000000000ll);
    int thr_i = 0x7fffffff;
    const int rmax = max(P.g[0], max(P.g[1], P.g[2]));
    for (int r = 0; r <= rmax; ++r) {
        const int z0 = (D > 2) ? max(qc[2] - r, 0) : 0, z1 = (D > 2) ? min(qc[2] + r, P.g[2] - 1) : 0;
        const int y0 = (D > 1) ? max(qc[1] - r, 0) : 0, y1 = (D > 1) ? min(qc[1] + r, P.g[1] - 1) : 0;
        const int x0 = max(qc[0] - r, 0), x1 = min(qc[0] + r, P.g[0] - 1);
        auto scan_cell = [&](int c) {
            const int s0 = __ldg(cstart + c), s1 = __ldg(cend + c);
            for (int s = s0; s < s1; ++s) {
                double d2 = 0.0;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const double df = __dsub_rn(q[j], __ldg(Xs + (int64_t)s * D + j));
                    d2 = __dadd_rn(d2, __dmul_rn(df, df));
                }
                const int ci = __ldg(ids_sorted + s);
                if (ci != qi && (d2 < thr || (d2 == thr && ci < thr_i))) {
                    double nd = d2; int ni = ci;
                    bool ins = false;
#pragma unroll
                    for (int w = 0; w < KMAX; ++w) {
                        if (w < k && (ins || nd < bd[w] || (nd == bd[w] && ni < bi[w]))) {
                            const double td = bd[w]; const int ti = bi[w];
                            bd[w] = nd; bi[w] = ni; nd = td; ni = ti; ins = true;
                        }
                    }
#pragma unroll
                    for (int w = 0; w < KMAX; ++w) if (w == k - 1) { thr = bd[w]; thr_i = bi[w]; }
                }
            }
        };
        for (int cz = z0; cz <= z1; ++cz)
            for (int cy = y0; cy <= y1; ++cy) {
                // on a face of the cube (|dz| == r or |dy| == r) every x cell belongs to the shell; otherwise only the
                // two cells with |dx| == r do
                const bool face = (D > 2 && (cz == qc[2] - r || cz == qc[2] + r)) || (D > 1 && (cy == qc[1] - r || cy == qc[1] + r));
                const int rowbase = (cz * P.g[1] + cy) * P.g[0];
                if (face) {
                    for (int cx = x0; cx <= x1; ++cx) scan_cell(rowbase + cx);
                } else {
                    if (qc[0] - r >= 0) scan_cell(rowbase + qc[0] - r);
                    if (r > 0 && qc[0] + r < P.g[0]) scan_cell(rowbase + qc[0] + r);
                }
            }
        // every point outside the scanned cube is at least r*h away
        const double reach = (double)r * P.h;
        if (thr < reach * reach * (1.0 - 1e-12)) break;
        if (x0 == 0 && x1 == P.g[0] - 1 && y0 == 0 && y1 == P.g[1] - 1 && z0 == 0 && z1 == P.g[2] - 1) break;   // whole grid scanned
    }
    const int64_t ql = qi - q_begin;
#pragma unroll
    for (int w = 0; w < KMAX; ++w)
        if (w < k) { out_idx[ql * k + w] = bi[w]; if (out_d2) out_d2[ql * k + w] = bd[w]; }
}

__global__ void grid_qpos_kernel(const int* __restrict__ ids_sorted, int n, int q_begin, int q_count, int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = ids_sorted[i];
    flags[i] = (id >= q_begin && id < q_begin + q_count) ? 1 : 0;
}

}  // namespace rvgp

using namespace rvgp;

static size_t al256(size_t b) { return (b + 255) / 256 * 256; }

struct GridWs { size_t cell, cell2, ids, ids2, xs, cstart, cend, flags, qpos, nsel, cub, cub_bytes, total; };

static GridWs grid_ws(int n, int D, int ncell) {
    GridWs w;
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, n);
    cub::DeviceSelect::Flagged(nullptr, b2, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, n);
    w.cub_bytes = b1 > b2 ? b1 : b2;
    size_t o = 0;
    w.cell = o; o += al256((size_t)n * 4);
    w.cell2 = o; o += al256((size_t)n * 4);
    w.ids = o; o += al256((size_t)n * 4);
    w.ids2 = o; o += al256((size_t)n * 4);
    w.xs = o; o += al256((size_t)n * D * 8);
    w.cstart = o; o += al256((size_t)ncell * 4);
    w.cend = o; o += al256((size_t)ncell * 4);
    w.flags = o; o += al256((size_t)n * 4);
    w.qpos = o; o += al256((size_t)n * 4);
    w.nsel = o; o += 256;
    w.cub = o; o += al256(w.cub_bytes);
    w.total = o;
    return w;
}

static int grid_dims(int n, int D, const double* lo, const double* hi, GridParams* P) {
    // ~2 cells per point of the bounding box volume (a surface leaves most of them empty), capped
    double vol = 1.0, ext[3] = {0, 0, 0};
    for (int j = 0; j < D; ++j) { ext[j] = hi[j] - lo[j]; if (!(ext[j] > 0)) ext[j] = 1e-300; vol *= ext[j]; }
    double target = 2.0 * n;
    if (target > 6.4e7) target = 6.4e7;
    double h = pow(vol / target, 1.0 / D);
    for (int it = 0; it < 60; ++it) {      // make sure the cell count fits
        double cells = 1.0;
        for (int j = 0; j < D; ++j) cells *= floor(ext[j] / h) + 1.0;
        if (cells <= 1.28e8) break;
        h *= 1.1;
    }
    P->D = D; P->h = h; P->inv_h = 1.0 / h;
    long long ncell = 1;
    for (int j = 0; j < 3; ++j) {
        P->lo[j] = j < D ? lo[j] : 0.0;
        P->g[j] = j < D ? (int)floor(ext[j] / h) + 1 : 1;
        ncell *= P->g[j];
    }
    return (int)ncell;
}

// lo/hi: HOST arrays (D) with the bounding box of X.
extern "C" int64_t rvgp_knn_grid_workspace_bytes(int n, int D, const double* lo, const double* hi) {
    GridParams P;
    const int ncell = grid_dims(n, D, lo, hi, &P);
    return (int64_t)grid_ws(n, D, ncell).total;
}

// Exact kNN through a uniform grid (D <= 3).  Same outputs as rvgp_knn_f64, bit-identical.
extern "C" int rvgp_knn_grid_f64(rvgp_handle_t hh, const double* X, int n, int D, const double* lo, const double* hi, int q_begin,
                                 int q_count, int k, int32_t* out_idx, double* out_d2, void* workspace, int64_t workspace_bytes) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, D >= 1 && D <= 3, "knn_grid: D must be 1, 2 or 3");
    RVGP_REQUIRE(h, k >= 1 && k <= 32 && k < n, "knn_grid: k must be in [1,32] and < n");
    RVGP_REQUIRE(h, q_begin >= 0 && q_count >= 0 && q_begin + q_count <= n, "knn_grid: bad query range");
    if (q_count == 0) return RVGP_OK;
    GridParams P;
    const int ncell = grid_dims(n, D, lo, hi, &P);
    GridWs w = grid_ws(n, D, ncell);
    if ((int64_t)w.total > workspace_bytes) return set_error(h, RVGP_ERR_CAPACITY, "knn_grid: workspace too small%s%s");
    char* base = (char*)workspace;
    int *cell = (int*)(base + w.cell), *cell2 = (int*)(base + w.cell2), *ids = (int*)(base + w.ids), *ids2 = (int*)(base + w.ids2);
    double* Xs = (double*)(base + w.xs);
    int *cstart = (int*)(base + w.cstart), *cend = (int*)(base + w.cend), *flags = (int*)(base + w.flags), *qpos = (int*)(base + w.qpos);
    int* nsel = (int*)(base + w.nsel);
    void* cubtmp = base + w.cub;
    grid_cell_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(X, n, P, cell, ids);
    RVGP_LAUNCH_OK(h, "grid_cell_kernel");
    int bits = 1;
    while ((1ll << bits) < ncell) ++bits;
    size_t cb = w.cub_bytes;
    RVGP_CUDA_OK(h, cub::DeviceRadixSort::SortPairs(cubtmp, cb, cell, cell2, ids, ids2, n, 0, bits, h->stream));
    RVGP_CUDA_OK(h, cudaMemsetAsync(cstart, 0, (size_t)ncell * 4, h->stream));
    RVGP_CUDA_OK(h, cudaMemsetAsync(cend, 0, (size_t)ncell * 4, h->stream));
    grid_bounds_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(cell2, n, cstart, cend);
    RVGP_LAUNCH_OK(h, "grid_bounds_kernel");
    grid_gather_kernel<<<cdiv((int64_t)n * D, 256), 256, 0, h->stream>>>(X, ids2, n, D, Xs);
    RVGP_LAUNCH_OK(h, "grid_gather_kernel");
    // sorted positions of this rank's queries
    grid_qpos_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(ids2, n, q_begin, q_count, flags);
    RVGP_LAUNCH_OK(h, "grid_qpos_kernel");
    cb = w.cub_bytes;
    RVGP_CUDA_OK(h, cub::DeviceSelect::Flagged(cubtmp, cb, cub::CountingInputIterator<int>(0), flags, qpos, nsel, n, h->stream));
    h->launches += 3;
#define RVGP_KG(DD)                                                                                                        \
    do {                                                                                                                   \
        if (k <= 16) knn_grid_kernel<DD, 16><<<cdiv(q_count, 128), 128, 0, h->stream>>>(Xs, ids2, cstart, cend, n, P, q_begin, q_count, k, qpos, out_idx, out_d2); \
        else knn_grid_kernel<DD, 32><<<cdiv(q_count, 128), 128, 0, h->stream>>>(Xs, ids2, cstart, cend, n, P, q_begin, q_count, k, qpos, out_idx, out_d2);         \
    } while (0)
    if (D == 1) RVGP_KG(1); else if (D == 2) RVGP_KG(2); else RVGP_KG(3);
#undef RVGP_KG
    RVGP_LAUNCH_OK(h, "knn_grid_kernel");
    return RVGP_OK;
}

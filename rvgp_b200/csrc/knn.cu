// K2: exact brute-force kNN (tiled pairwise distances, shared-memory staging, register top-k).
//
// Replaces sklearn.neighbors.kneighbors_graph(X, nb, mode='connectivity', include_self=False)
// (reference RVGP/geometry.py:103-110).  Bit-exact neighbour indices need the same arithmetic as
// sklearn's KD-tree "rdist": sum_j (x_j - y_j)^2 accumulated coordinate by coordinate in FP64 WITHOUT
// FMA contraction (__dsub_rn/__dmul_rn/__dadd_rn), the sample itself removed by INDEX, and exact ties
// broken by lowest index.
//
// One thread owns QPT queries and keeps their k best (distance, index) pairs sorted in registers;
// candidate tiles are staged in shared memory and read by broadcast.  Inserts are rare after the first
// tiles (expected k*ln(n/k) per query), so the steady state is 3D-2 FP64 ops + one compare per pair.
// Rooflines: FP64 pipe, (3D-1) n^2 flop; streamed candidate bytes (n/Tq) n D 8.
#include "common.cuh"

namespace rvgp {

constexpr int KNN_THREADS = 128;

template <int D, int KMAX, int QPT>
__global__ void __launch_bounds__(KNN_THREADS)
knn_kernel(const double* __restrict__ X, int n, int dreal, int q_begin, int q_count, int k,
           int* __restrict__ out_idx, double* __restrict__ out_d2) {
    constexpr int KNN_TILE = (D <= 8) ? 512 : 4096 / D;   // candidates per shared-memory tile (<= 32 KB)
    __shared__ double cand[KNN_TILE * D];
    const int t = threadIdx.x;
    int qid[QPT];
    double q[QPT][D];
    double bd[QPT][KMAX];
    int bi[QPT][KMAX];
    double thr[QPT];
#pragma unroll
    for (int s = 0; s < QPT; ++s) {
        const int ql = (blockIdx.x * QPT + s) * KNN_THREADS + t;   // local query id
        qid[s] = (ql < q_count) ? q_begin + ql : -1;
#pragma unroll
        for (int j = 0; j < D; ++j) q[s][j] = (qid[s] >= 0 && j < dreal) ? __ldg(X + (int64_t)qid[s] * dreal + j) : 0.0;
#pragma unroll
        for (int r = 0; r < KMAX; ++r) { bd[s][r] = __longlong_as_double(0x7ff0000000000000ll); bi[s][r] = 0x7fffffff; }
        thr[s] = __longlong_as_double(0x7ff0000000000000ll);
    }

    for (int c0 = 0; c0 < n; c0 += KNN_TILE) {
        const int tile = min(KNN_TILE, n - c0);
        __syncthreads();
        for (int e = t; e < tile * D; e += KNN_THREADS) {
            const int c = e / D, j = e % D;
            cand[e] = (j < dreal) ? __ldg(X + (int64_t)(c0 + c) * dreal + j) : 0.0;
        }
        __syncthreads();
        for (int c = 0; c < tile; ++c) {
            double cj[D];
#pragma unroll
            for (int j = 0; j < D; ++j) cj[j] = cand[c * D + j];
#pragma unroll
            for (int s = 0; s < QPT; ++s) {
                double d2 = 0.0;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const double df = __dsub_rn(q[s][j], cj[j]);
                    d2 = __dadd_rn(d2, __dmul_rn(df, df));
                }
                // candidates are visited in ascending index order, so "strictly smaller" keeps the lowest
                // index among exact ties
                if (d2 < thr[s] && (c0 + c) != qid[s] && qid[s] >= 0) {
                    double nd = d2;
                    int ni = c0 + c;
                    bool ins = false;   // once the slot is found every later entry shifts down by one
#pragma unroll
                    for (int r = 0; r < KMAX; ++r) {
                        if (r < k && (ins || nd < bd[s][r])) {
                            const double td = bd[s][r]; const int ti = bi[s][r];
                            bd[s][r] = nd; bi[s][r] = ni; nd = td; ni = ti;
                            ins = true;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < KMAX; ++r) if (r == k - 1) thr[s] = bd[s][r];
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < QPT; ++s) {
        if (qid[s] < 0) continue;
        const int64_t ql = qid[s] - q_begin;
#pragma unroll
        for (int r = 0; r < KMAX; ++r) {
            if (r < k) {
                out_idx[ql * k + r] = bi[s][r];
                if (out_d2) out_d2[ql * k + r] = bd[s][r];
            }
        }
    }
}

template <int D, int KMAX>
static void launch_knn(Handle* h, const double* X, int n, int dreal, int q_begin, int q_count, int k, int* out_idx,
                       double* out_d2) {
    constexpr int QPT = (D <= 4 && KMAX <= 16) ? 2 : 1;
    const int grid = cdiv(q_count, KNN_THREADS * QPT);
    knn_kernel<D, KMAX, QPT><<<grid, KNN_THREADS, 0, h->stream>>>(X, n, dreal, q_begin, q_count, k, out_idx, out_d2);
}

}  // namespace rvgp

using namespace rvgp;

// X: (n, D) row-major FP64 (all candidates).  Queries are rows [q_begin, q_begin + q_count) (row sharding for
// multi-GPU: every rank holds all of X and answers its own query range).  out_idx: (q_count, k) int32,
// ascending (distance, index); out_d2 (nullable): the squared distances.
extern "C" int rvgp_knn_f64(rvgp_handle_t hh, const double* X, int n, int D, int q_begin, int q_count, int k,
                            int32_t* out_idx, double* out_d2) {
    Handle* h = H(hh);
    RVGP_REQUIRE(h, n >= 1 && D >= 1 && D <= 64, "knn: D must be in [1,64]");
    RVGP_REQUIRE(h, k >= 1 && k <= 32 && k < n, "knn: k must be in [1,32] and < n");
    RVGP_REQUIRE(h, q_begin >= 0 && q_count >= 0 && q_begin + q_count <= n, "knn: bad query range");
    if (q_count == 0) return RVGP_OK;
#define RVGP_KNN(DP)                                                                                         \
    do {                                                                                                     \
        if (k <= 16) launch_knn<DP, 16>(h, X, n, D, q_begin, q_count, k, out_idx, out_d2);                   \
        else launch_knn<DP, 32>(h, X, n, D, q_begin, q_count, k, out_idx, out_d2);                           \
    } while (0)
    // D is padded up to the next instantiated size with zero coordinates: (0-0)^2 adds +0.0 exactly
    if (D <= 2) RVGP_KNN(2);
    else if (D <= 3) RVGP_KNN(3);
    else if (D <= 4) RVGP_KNN(4);
    else if (D <= 6) RVGP_KNN(6);
    else if (D <= 8) RVGP_KNN(8);
    else if (D <= 12) RVGP_KNN(12);
    else if (D <= 16) RVGP_KNN(16);
    else if (D <= 24) RVGP_KNN(24);
    else if (D <= 32) RVGP_KNN(32);
    else RVGP_KNN(64);
#undef RVGP_KNN
    RVGP_LAUNCH_OK(h, "knn_kernel");
    return RVGP_OK;
}

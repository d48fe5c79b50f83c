// K4: geodesic neighbourhoods = literal emulation of the reference's Fibonacci-heap Dijkstra.
//
// Replaces the per-source loop of _geodesic_neigborhood_tangents (reference
// RVGP/lib/ptu_dijkstra.pyx:361-394) and its heap (pyx:444-690).  All edge weights are 1.0, so WHICH
// K+1 nodes form a neighbourhood is decided purely by the heap's pop order among equal keys; the heap is
// therefore reproduced operation by operation (insert :579-589, remove_min :642-690, link :614-636,
// add_child :507-520, add_sibling :526-536, remove :542-560) in the ORIGINAL node numbering.
//
// One thread per source.  A source touches at most 1 + K*maxdeg nodes, so each thread owns a small node
// pool (16-bit links) and an open-addressing hash map "global id -> pool slot" in a private slab of the
// caller's workspace.  Integer, divergent, latency-bound work; n independent sources.
// With unit weights decrease_val (pyx:595-608) can never fire (a node's first discovery is through its
// earliest-popped neighbour); the kernel checks that and raises a flag instead of implementing it.
#include "common.cuh"

namespace rvgp {

constexpr unsigned short NIL = 0xFFFF;
enum : unsigned char { ST_SCANNED = 0, ST_IN_HEAP = 2 };

struct Pool {
    int* id;
    unsigned short *parent, *left, *right, *child, *hpos;
    unsigned char *val, *rank, *state;
    int* hkey;
    unsigned short* hslot;
    int hmask, hshift;
};

__device__ __forceinline__ unsigned short rightmost(const Pool& P, unsigned short n) {
    while (P.right[n] != NIL) n = P.right[n];
    return n;
}
__device__ __forceinline__ unsigned short leftmost(const Pool& P, unsigned short n) {
    while (P.left[n] != NIL) n = P.left[n];
    return n;
}
__device__ __forceinline__ void add_sibling(const Pool& P, unsigned short node, unsigned short sib) {   // pyx:526-536
    const unsigned short t = rightmost(P, node);
    P.right[t] = sib;
    P.left[sib] = t;
    P.right[sib] = NIL;
    P.parent[sib] = P.parent[node];
    if (P.parent[sib] != NIL) P.rank[P.parent[sib]] += 1;
}
__device__ __forceinline__ void add_child(const Pool& P, unsigned short node, unsigned short ch) {      // pyx:507-520
    P.parent[ch] = node;
    if (P.child[node] != NIL) {
        add_sibling(P, P.child[node], ch);
    } else {
        P.child[node] = ch;
        P.right[ch] = NIL;
        P.left[ch] = NIL;
        P.rank[node] = 1;
    }
}
__device__ __forceinline__ void remove_node(const Pool& P, unsigned short node) {                       // pyx:542-560
    const unsigned short p = P.parent[node];
    if (p != NIL) {
        P.rank[p] -= 1;
        if (P.left[node] != NIL) P.child[p] = P.left[node];
        else if (P.right[node] != NIL) P.child[p] = P.right[node];
        else P.child[p] = NIL;
    }
    if (P.left[node] != NIL) P.right[P.left[node]] = P.right[node];
    if (P.right[node] != NIL) P.left[P.right[node]] = P.left[node];
    P.left[node] = NIL;
    P.right[node] = NIL;
    P.parent[node] = NIL;
}

__global__ void __launch_bounds__(128)
geodesic_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, int n, int K, int cap, int hcap,
                unsigned char* __restrict__ ws, int64_t slab_bytes, int* __restrict__ seq, int* __restrict__ counts,
                int* __restrict__ flags, int src_begin, int src_end) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const int Kp1 = K + 1;
    // carve this thread's slab
    unsigned char* base = ws + (int64_t)tid * slab_bytes;
    Pool P;
    P.id = (int*)base;                       base += (int64_t)cap * 4;
    P.hkey = (int*)base;                     base += (int64_t)hcap * 4;
    P.parent = (unsigned short*)base;        base += (int64_t)cap * 2;
    P.left = (unsigned short*)base;          base += (int64_t)cap * 2;
    P.right = (unsigned short*)base;         base += (int64_t)cap * 2;
    P.child = (unsigned short*)base;         base += (int64_t)cap * 2;
    P.hpos = (unsigned short*)base;          base += (int64_t)cap * 2;
    P.hslot = (unsigned short*)base;         base += (int64_t)hcap * 2;
    P.val = base;                            base += cap;
    P.rank = base;                           base += cap;
    P.state = base;
    P.hmask = hcap - 1;
    int lg = 0;
    while ((1 << lg) < hcap) ++lg;
    P.hshift = 32 - lg;
    for (int e = 0; e < hcap; ++e) P.hkey[e] = -1;

    unsigned short roots[64];   // roots_by_rank (pyx:573); ranks stay < log2(cap)

    for (int src = src_begin + tid; src < src_end; src += nthreads) {
        int count = 0;            // nodes in the pool
        unsigned short min_node = NIL;
        int scanned = 0;
        bool overflow = false;

        auto new_node = [&](int gid, unsigned char v) -> unsigned short {   // initialize_node pyx:465-479
            const unsigned short s = (unsigned short)count++;
            P.id[s] = gid; P.val[s] = v; P.rank[s] = 0; P.state[s] = ST_IN_HEAP;
            P.parent[s] = NIL; P.left[s] = NIL; P.right[s] = NIL; P.child[s] = NIL;
            unsigned int hp = ((unsigned int)gid * 2654435761u) >> P.hshift;
            while (P.hkey[hp] != -1) hp = (hp + 1) & P.hmask;
            P.hkey[hp] = gid; P.hslot[hp] = s; P.hpos[s] = (unsigned short)hp;
            return s;
        };
        auto insert_node = [&](unsigned short node) {                        // pyx:579-589
            if (min_node != NIL) {
                add_sibling(P, min_node, node);
                if (P.val[node] < P.val[min_node]) min_node = node;
            } else {
                min_node = node;
            }
        };

        insert_node(new_node(src, 0));

        while (min_node != NIL && scanned <= K) {
            // ---- remove_min (pyx:642-690) ----
            unsigned short temp, temp_right, out;
            const unsigned short mn = min_node;
            if (P.child[mn] != NIL) {
                temp = leftmost(P, P.child[mn]);
                while (temp != NIL) {
                    temp_right = P.right[temp];
                    remove_node(P, temp);
                    add_sibling(P, mn, temp);
                    temp = temp_right;
                }
                P.child[mn] = NIL;
            }
            temp = leftmost(P, mn);
            bool emptied = false;
            if (temp == mn) {
                if (P.right[mn] != NIL) temp = P.right[mn];
                else { emptied = true; }
            }
            out = mn;
            if (emptied) {
                min_node = NIL;
            } else {
                remove_node(P, mn);
                min_node = temp;
                for (int r = 0; r < 64; ++r) roots[r] = NIL;
                while (temp != NIL) {
                    if (P.val[temp] < P.val[min_node]) min_node = temp;
                    temp_right = P.right[temp];
                    // ---- link (pyx:614-636), tail recursion as a loop ----
                    unsigned short node = temp;
                    for (;;) {
                        const int r = P.rank[node];
                        if (roots[r] == NIL) { roots[r] = node; break; }
                        const unsigned short ln = roots[r];
                        roots[r] = NIL;
                        if (P.val[node] < P.val[ln] || node == min_node) {
                            remove_node(P, ln);
                            add_child(P, node, ln);
                        } else {
                            remove_node(P, node);
                            add_child(P, ln, node);
                            node = ln;
                        }
                    }
                    temp = temp_right;
                }
            }
            // ---- driver (pyx:378-394) ----
            const unsigned short v = out;
            P.state[v] = ST_SCANNED;
            const int j = P.id[v];
            seq[(int64_t)src * Kp1 + scanned] = j;
            scanned += 1;
            if (scanned <= K) {
                const unsigned char nv = P.val[v] + 1;
                const int e1 = __ldg(indptr + j + 1);
                for (int e = __ldg(indptr + j); e < e1; ++e) {
                    const int c = __ldg(indices + e);
                    unsigned int hp = ((unsigned int)c * 2654435761u) >> P.hshift;
                    int slot = -1;
                    while (P.hkey[hp] != -1) {
                        if (P.hkey[hp] == c) { slot = P.hslot[hp]; break; }
                        hp = (hp + 1) & P.hmask;
                    }
                    if (slot < 0) {                       // NOT_IN_HEAP
                        if (count >= cap) { overflow = true; continue; }
                        insert_node(new_node(c, nv));
                    } else if (P.state[slot] != ST_SCANNED && P.val[slot] > nv) {
                        atomicOr(flags, 2);               // decrease_val would fire: impossible with unit weights
                    }
                }
            }
        }
        counts[src] = scanned;
        if (scanned < Kp1) atomicOr(flags, 1);            // short component: stale tail (fixed up afterwards)
        if (overflow) atomicOr(flags, 4);
        for (int s = 0; s < count; ++s) P.hkey[P.hpos[s]] = -1;
    }
}

// pyx:350: the index buffer is allocated once outside the source loop, so a source whose component has
// fewer than K+1 nodes keeps the PREVIOUS source's entries in the tail.  Sequential by construction.
__global__ void geodesic_fix_stale_kernel(int n, int Kp1, int* __restrict__ seq, const int* __restrict__ counts,
                                          const int* __restrict__ flags) {
    if (!(*flags & 1)) return;
    for (int i = 0; i < n; ++i) {
        const int c = counts[i];
        for (int q = c; q < Kp1; ++q) seq[(int64_t)i * Kp1 + q] = (i > 0) ? seq[(int64_t)(i - 1) * Kp1 + q] : 0;
    }
}

static void geodesic_dims(int K, int maxdeg, int* cap, int* hcap, int64_t* slab) {
    long long c = 1 + (long long)K * maxdeg;
    if (c > 60000) c = 60000;
    int hc = 64;
    while (hc < 2 * c) hc <<= 1;
    *cap = (int)c;
    *hcap = hc;
    int64_t b = (int64_t)c * 4 + (int64_t)hc * 4 + (int64_t)c * 2 * 5 + (int64_t)hc * 2 + (int64_t)c * 3;
    *slab = (b + 15) / 16 * 16;
}

static int geodesic_threads(Handle* h, int n, int64_t slab) {
    int64_t t = (int64_t)(h ? h->sm_count : 148) * 512;
    const int64_t budget = (int64_t)2 << 30;   // keep the slab array under 2 GiB
    if (t * slab > budget) t = budget / slab;
    if (t > n) t = n;
    t = (t + 127) / 128 * 128;
    return (int)(t < 128 ? 128 : t);
}

}  // namespace rvgp

using namespace rvgp;

// maxdeg: maximum number of stored entries in a CSR row (self loop included is fine).
extern "C" int64_t rvgp_geodesic_workspace_bytes(rvgp_handle_t hh, int n, int K, int maxdeg) {
    int cap, hcap; int64_t slab;
    geodesic_dims(K, maxdeg, &cap, &hcap, &slab);
    return (int64_t)geodesic_threads(H(hh), n, slab) * slab;
}

// seq: (n, K+1) int32 popped node ids in pop order; counts: (n) pops per source; flags: device int32
// (bit0 short component seen, bit1 decrease_val would have fired, bit2 pool overflow) -- zeroed here.
// The _range form computes only the sources [src_begin, src_begin + src_count) (rows of the FULL-size seq / counts arrays) and
// leaves the stale-tail pass to the caller: sources are independent, so a multi-GPU caller gives every rank a contiguous
// range, all-gathers the rows and runs rvgp_geodesic_fix_stale once on the complete arrays (the stale tail of a short
// component copies from the PREVIOUS source's row, pyx:350, hence needs them all).
static int geodesic_run(Handle* h, const int32_t* indptr, const int32_t* indices, int n, int K, int maxdeg, int src_begin,
                        int src_count, int32_t* seq, int32_t* counts, int32_t* flags, void* workspace, int64_t workspace_bytes) {
    RVGP_REQUIRE(h, n >= 1 && K >= 1 && K < 250, "geodesic: K must be in [1,250)");
    RVGP_REQUIRE(h, K < n, "Geodesic neighborhood size must be less than the total number of samples");
    RVGP_REQUIRE(h, src_begin >= 0 && src_count >= 0 && src_begin + src_count <= n, "geodesic: source range outside [0, n)");
    int cap, hcap; int64_t slab;
    geodesic_dims(K, maxdeg, &cap, &hcap, &slab);
    const int nthreads = geodesic_threads(h, n, slab);
    if ((int64_t)nthreads * slab > workspace_bytes) return set_error(h, RVGP_ERR_CAPACITY, "geodesic: workspace too small%s%s");
    RVGP_CUDA_OK(h, cudaMemsetAsync(flags, 0, sizeof(int), h->stream));
    if (src_count == 0) return RVGP_OK;
    geodesic_kernel<<<nthreads / 128, 128, 0, h->stream>>>(indptr, indices, n, K, cap, hcap, (unsigned char*)workspace, slab,
                                                           seq, counts, flags, src_begin, src_begin + src_count);
    RVGP_LAUNCH_OK(h, "geodesic_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_geodesic_fix_stale(rvgp_handle_t hh, int n, int K, int32_t* seq, const int32_t* counts, const int32_t* flags) {
    Handle* h = H(hh);
    geodesic_fix_stale_kernel<<<1, 1, 0, h->stream>>>(n, K + 1, seq, counts, flags);
    RVGP_LAUNCH_OK(h, "geodesic_fix_stale_kernel");
    return RVGP_OK;
}

extern "C" int rvgp_geodesic_neighbourhoods_range(rvgp_handle_t hh, const int32_t* indptr, const int32_t* indices, int n, int K,
                                                  int maxdeg, int src_begin, int src_count, int32_t* seq, int32_t* counts,
                                                  int32_t* flags, void* workspace, int64_t workspace_bytes) {
    return geodesic_run(H(hh), indptr, indices, n, K, maxdeg, src_begin, src_count, seq, counts, flags, workspace, workspace_bytes);
}

extern "C" int rvgp_geodesic_neighbourhoods(rvgp_handle_t hh, const int32_t* indptr, const int32_t* indices, int n, int K,
                                            int maxdeg, int32_t* seq, int32_t* counts, int32_t* flags, void* workspace,
                                            int64_t workspace_bytes) {
    int rc = geodesic_run(H(hh), indptr, indices, n, K, maxdeg, 0, n, seq, counts, flags, workspace, workspace_bytes);
    if (rc) return rc;
    return rvgp_geodesic_fix_stale(hh, n, K, seq, counts, flags);
}
